"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own modules (mmnas/model/{modules,mixed,full_vqa,hygr_vqa}.py,
mmnas/utils/ops_adapter.py, and the body of relation_embedding from
mmnas/loader/load_data_vqa.py:7-33, whose module cannot be imported because it needs spaCy),
runs them on seeded inputs in float32 on CPU and stores inputs, weights, outputs and gradients
as .npz.  The vectors pin oracle/mmnas_oracle.py (tests/test_oracle_golden.py) and are a
second, oracle-independent anchor for the CUDA path (tests/test_gpu_golden.py).
"""
import ast
import os
import sys

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def op_cfg(h):
    return Bag(HSIZE=h, DROPOUT_R=0.0, REL_SIZE=64, OPS_NORM=True, OPS_RESIDUAL=True)


def net_cfg(h, genotype=None, nodes=None):
    return Bag(HSIZE=h, DROPOUT_R=0.0, REL_SIZE=64, OPS_NORM=True, OPS_RESIDUAL=True, LAYERS=1,
               BBOX_FEATURE=False, FRCNFEAT_SIZE=32, BBOXFEAT_EMB_SIZE=32, WORD_EMBED_SIZE=16,
               ATTFLAT_GLIMPSES=1, ATTFLAT_OUT_SIZE=2 * h, ATTFLAT_MLP_SIZE=48,
               GENOTYPE=genotype, NODES=nodes, ALPHA_INIT_TYPE='normal')


def npify(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()
            if v is not None}


def masks(b, n, lens):
    m = torch.zeros(b, 1, 1, n, dtype=torch.bool)
    for i, l in enumerate(lens):
        m[i, 0, 0, l:] = True
    return m


def golden_ops(mods):
    """Each of the four candidate blocks, fwd + bwd, H=128 (2 heads), ragged masks incl. a fully padded sample."""
    out = {}
    h, b, nx, ny = 128, 3, 7, 5
    for name in ('self_att_64', 'rel_self_att_64', 'guided_att_64', 'feed_forward'):
        torch.manual_seed(888)
        op = mods['OpsAdapter']().OPS[name](op_cfg(h), True, True)
        with torch.no_grad():                       # make LN affine / biases non-trivial
            for n_, p_ in op.named_parameters():
                if n_.endswith('a_2'):
                    p_.add_(0.1 * torch.randn_like(p_))
                if n_.endswith('b_2') or n_.endswith('bias'):
                    p_.add_(0.1 * torch.randn_like(p_))
        x = torch.randn(b, nx, h, requires_grad=True)
        y = torch.randn(b, ny, h, requires_grad=True)
        rel = torch.relu(torch.randn(b, nx, nx, 64)).requires_grad_(True)
        x_mask = masks(b, nx, [7, 4, 0])            # sample 2: every key padded -> uniform attention row
        y_mask = masks(b, ny, [5, 2, 0])
        o = op(x, y, x_mask, y_mask, rel)
        go = torch.randn_like(o)
        o.backward(go)
        rec = {'x': x, 'y': y, 'rel': rel, 'x_mask': x_mask, 'y_mask': y_mask, 'out': o, 'gout': go,
               'gx': x.grad}
        if y.grad is not None:
            rec['gy'] = y.grad
        if rel.grad is not None:
            rec['grel'] = rel.grad
        for n_, p_ in op.named_parameters():
            rec['p.' + n_] = p_
            rec['g.' + n_] = p_.grad
        out.update({'%s/%s' % (name, k): v for k, v in npify(rec).items()})
    np.savez_compressed(os.path.join(OUT, 'ops_h128.npz'), **out)


def golden_mixed(mods):
    """MixedOp 'full' mode on the decoder candidate set: output, alpha_gate.grad, alpha_prob.grad."""
    MixedOp = mods['MixedOp']
    h, b, nx, ny = 64, 2, 6, 4
    torch.manual_seed(888)
    m = MixedOp(op_cfg(h), 'dec_safe')
    with torch.no_grad():
        m.alpha_prob.copy_(torch.tensor([-1., 1., -1., 0.5]))
    MixedOp.MODE = 'full'
    m.binarize()
    x = torch.randn(b, nx, h, requires_grad=True)
    y = torch.randn(b, ny, h, requires_grad=True)
    rel = torch.relu(torch.randn(b, nx, nx, 64))
    x_mask, y_mask = masks(b, nx, [6, 3]), masks(b, ny, [4, 1])
    o = m(x, y, x_mask, y_mask, rel)
    go = torch.randn_like(o)
    o.backward(go)
    gate_grad = m.alpha_gate.grad.clone()
    m.set_arch_param_grad()
    rec = {'x': x, 'y': y, 'rel': rel, 'x_mask': x_mask, 'y_mask': y_mask, 'out': o, 'gout': go,
           'gx': x.grad, 'gy': y.grad, 'active': np.array(m.active_index), 'inactive': np.array(m.inactive_index),
           'gate_grad': gate_grad, 'prob_grad': m.alpha_prob.grad}
    for n_, p_ in m.named_parameters():
        rec['p.' + n_] = p_
        if p_.grad is not None and 'alpha' not in n_:
            rec['g.' + n_] = p_.grad
    MixedOp.MODE = None
    np.savez_compressed(os.path.join(OUT, 'mixed_h64.npz'), **npify(rec))


def synth_inputs(b, ny, nx, feat, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    frcn = torch.relu(torch.randn(b, ny, feat, generator=g))
    n_obj = [ny, max(2, ny // 2)] + [ny] * (b - 2)
    rel = torch.zeros(b, ny, ny, 4)
    for i in range(b):
        frcn[i, n_obj[i]:] = 0
        rel[i, :n_obj[i], :n_obj[i]] = torch.randn(n_obj[i], n_obj[i], 4, generator=g)
    ques = torch.randint(3, vocab, (b, nx), generator=g)
    ques[0, nx - 2:] = 0
    bbox = torch.zeros(b, ny, 5)
    rel_q = torch.zeros(b, nx, nx, 3)
    return frcn, bbox, rel, ques, rel_q


def golden_net_full(mods):
    """Tiny Net_Full-VQA train step (train_vqa.py:294-299): loss, logits, every parameter gradient."""
    geno = {'enc': [['self_att_64'], ['feed_forward']],
            'dec': [['guided_att_64'], ['rel_self_att_64'], ['self_att_64'], ['feed_forward']]}
    h, b, ny, nx, vocab, ans = 64, 2, 6, 5, 30, 11
    torch.manual_seed(888)
    np.random.seed(888)
    init = {'token_size': vocab, 'ans_size': ans,
            'pretrained_emb': (0.1 * np.random.randn(vocab, 16)).astype(np.float32)}
    net = mods['Net_Full'](net_cfg(h, genotype=geno), init)
    inputs = synth_inputs(b, ny, nx, 32, vocab, 1)
    target = torch.rand(b, ans, generator=torch.Generator().manual_seed(2)).round()
    pred = net(inputs)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, target, reduction='sum')
    loss = loss + 0 * sum(p.sum() for p in net.parameters())
    loss.backward()
    rec = {'frcn': inputs[0], 'bbox': inputs[1], 'rel': inputs[2], 'ques': inputs[3], 'rel_q': inputs[4],
           'target': target, 'pred': pred, 'loss': loss, 'genotype': np.array(repr(geno))}
    for n_, p_ in net.named_parameters():
        rec['p.' + n_] = p_
        rec['g.' + n_] = p_.grad
    np.savez_compressed(os.path.join(OUT, 'net_full_h64.npz'), **npify(rec))


def golden_net_full_task(task):
    """Tiny Net_Full of full_vgd.py / full_itm.py (VGD: 7-token queries, every region valid, KLD log-softmax scores +
    box regression; ITM: sigmoid matching score): outputs and every gradient of a fixed linear functional of them."""
    import importlib
    Net_Full = importlib.import_module('mmnas.model.full_' + task).Net_Full
    geno = {'enc': [['self_att_64'], ['feed_forward']],
            'dec': [['guided_att_64'], ['rel_self_att_64'], ['guided_att_64'], ['feed_forward']]}
    h, b, ny, nx, vocab = 64, 3, 6, 7, 30
    torch.manual_seed(888)
    np.random.seed(888)
    init = {'token_size': vocab, 'ans_size': 5, 'pretrained_emb': (0.1 * np.random.randn(vocab, 16)).astype(np.float32)}
    cfg = net_cfg(h, genotype=geno)
    cfg.SCORES_LOSS = 'kld'
    net = Net_Full(cfg, init)
    inputs = synth_inputs(b, ny, nx, 32, vocab, 7)
    g = torch.Generator().manual_seed(8)
    outs = net(inputs)
    outs = outs if isinstance(outs, tuple) else (outs,)
    ws = [torch.randn(o.shape, generator=g) for o in outs]
    sum((o * w).sum() for o, w in zip(outs, ws)).backward()
    rec = {'frcn': inputs[0], 'bbox': inputs[1], 'rel': inputs[2], 'ques': inputs[3], 'rel_q': inputs[4],
           'genotype': np.array(repr(geno))}
    for i, (o, w) in enumerate(zip(outs, ws)):
        rec['out%d' % i] = o
        rec['w%d' % i] = w
    for n_, p_ in net.named_parameters():
        rec['p.' + n_] = p_
        rec['g.' + n_] = p_.grad
    np.savez_compressed(os.path.join(OUT, 'net_full_%s_h64.npz' % task), **npify(rec))


def golden_net_search(mods):
    """Tiny Net_Search-VQA arch step (search_vqa.py:317-332, MODE='full'): loss, alpha_gate grads,
    alpha_prob grads after set_arch_param_grad, alpha_prob after one alpha Adam step, genotype."""
    MixedOp = mods['MixedOp']
    h, b, ny, nx, vocab, ans = 64, 2, 6, 5, 30, 11
    torch.manual_seed(888)
    np.random.seed(888)
    init = {'token_size': vocab, 'ans_size': ans,
            'pretrained_emb': (0.1 * np.random.randn(vocab, 16)).astype(np.float32)}
    net = mods['Net_Search'](net_cfg(h, nodes={'enc': 2, 'dec': 3}), init)
    # init_arch (hygr_vqa.py:142-156) hard-codes "first 12 alphas are encoder nodes"; with 2+3 nodes it
    # leaves decoder alphas with 2 entries.  Overwrite the alpha VALUES (state, not code) with seeded ones.
    ga = torch.Generator().manual_seed(6)
    for n_, p_ in net.named_alpha_prob_parameters():
        p_.data = torch.randn(2 if 'cells_enc' in n_ else 4, generator=ga)
    alpha_optim = torch.optim.Adam(net.alpha_prob_parameters(), 0.1, betas=(0., 0.999), weight_decay=0)
    inputs = synth_inputs(b, ny, nx, 32, vocab, 3)
    target = torch.rand(b, ans, generator=torch.Generator().manual_seed(4)).round()
    MixedOp.MODE = 'full'
    net.reset_binary_gates()
    net.unused_modules_off()
    choices = {'enc': [], 'dec': []}
    for n_, m in net.named_modules():
        if str(m).startswith('MixedOp'):
            choices['enc' if 'cells_enc' in n_ else 'dec'].append(m.active_index[0])
    pred = net(inputs)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, target, reduction='sum')
    loss = loss + 0 * sum(p.sum() for p in net.alpha_prob_parameters())
    loss = loss + 0 * sum(p.sum() for p in net.net_parameters())
    net.zero_grad()
    loss.backward()
    rec = {'frcn': inputs[0], 'bbox': inputs[1], 'rel': inputs[2], 'ques': inputs[3], 'rel_q': inputs[4],
           'target': target, 'pred': pred, 'loss': loss,
           'choices_enc': np.array(choices['enc']), 'choices_dec': np.array(choices['dec'])}
    for n_, p_ in net.named_parameters():
        rec['p.' + n_] = p_.detach().clone()
        if 'alpha_gate' in n_:
            rec['g.' + n_] = p_.grad.clone()
    net.set_arch_param_grad()
    for n_, p_ in net.named_alpha_prob_parameters():
        rec['g.' + n_] = p_.grad.clone()
    alpha_optim.step()
    for n_, p_ in net.named_alpha_prob_parameters():
        rec['after.' + n_] = p_.detach().clone()
    net.unused_modules_back()
    MixedOp.MODE = None
    rec['genotype'] = np.array(repr(net.genotype()))
    np.savez_compressed(os.path.join(OUT, 'net_search_h64.npz'), **npify(rec))


def golden_mixed_two(mods):
    """MixedOp in MODE 'two' (mixed.py:136-148, :179-191, :200-208): the pair binarize() draws under seed 888, the
    gated two-candidate output, alpha_gate.grad, the 2 x 2 alpha_prob.grad rule, one alpha Adam step (lr 0.1,
    betas (0, .999): search_vqa.py:190-197) and the logsumexp-preserving rescale."""
    MixedOp = mods['MixedOp']
    h, b, nx, ny = 64, 2, 6, 4
    torch.manual_seed(888)
    m = MixedOp(op_cfg(h), 'dec_safe')
    with torch.no_grad():
        m.alpha_prob.copy_(torch.tensor([0.3, 1., -0.5, 0.5]))
    opt = torch.optim.Adam([m.alpha_prob], 0.1, betas=(0., 0.999), weight_decay=0)
    MixedOp.MODE = 'two'
    torch.manual_seed(888)
    m.binarize()
    active, inactive = list(m.active_index), list(m.inactive_index)
    x = torch.randn(b, nx, h, requires_grad=True)
    y = torch.randn(b, ny, h, requires_grad=True)
    rel = torch.relu(torch.randn(b, nx, nx, 64))
    x_mask, y_mask = masks(b, nx, [6, 3]), masks(b, ny, [4, 1])
    saved = {i: m.candidate_ops[i] for i in range(m.n_choices) if i not in active + inactive}
    for i in saved:                                # Net_Search.unused_modules_off (hygr_vqa.py:175-187)
        m.candidate_ops[i] = None
    o = m(x, y, x_mask, y_mask, rel)
    go = torch.randn_like(o)
    o.backward(go)
    for i, op in saved.items():
        m.candidate_ops[i] = op
    gate_grad = m.alpha_gate.grad.clone()
    alpha_before = m.alpha_prob.detach().clone()
    m.set_arch_param_grad()
    prob_grad = m.alpha_prob.grad.clone()
    opt.step()
    alpha_adam = m.alpha_prob.detach().clone()
    m.rescale_updated_arch_param()
    rec = {'x': x, 'y': y, 'rel': rel, 'x_mask': x_mask, 'y_mask': y_mask, 'out': o, 'gout': go, 'gx': x.grad,
           'gy': y.grad, 'active': np.array(active), 'inactive': np.array(inactive), 'gate_grad': gate_grad,
           'prob_grad': prob_grad, 'alpha_before': alpha_before, 'alpha_adam': alpha_adam,
           'alpha_rescaled': m.alpha_prob.detach().clone()}
    for n_, p_ in m.named_parameters():
        if 'alpha_prob' in n_:
            rec['p.' + n_] = alpha_before
        elif 'alpha_gate' in n_:
            rec['p.' + n_] = p_
        else:
            rec['p.' + n_] = p_
            if p_.grad is not None:
                rec['g.' + n_] = p_.grad
    MixedOp.MODE = None
    np.savez_compressed(os.path.join(OUT, 'mixed_two_h64.npz'), **npify(rec))


def golden_sampling(mods):
    """The architecture samples the reference draws under its search seed: Net_Search.reset_binary_gates
    (hygr_vqa.py:168-172 over MixedOp.binarize mixed.py:150-156) on the 12 + 18 node supernet with the init_arch prior
    (hygr_vqa.py:142-156), torch.manual_seed(888) (search_vqa.py:62), CPU generator, five consecutive draws; plus the
    draws after one alpha step moved the distribution."""
    MixedOp = mods['MixedOp']
    np.random.seed(888)
    init = {'token_size': 30, 'ans_size': 11, 'pretrained_emb': np.zeros((30, 16), np.float32)}
    cfg = net_cfg(64, nodes={'enc': 12, 'dec': 18})
    torch.manual_seed(888)
    net = mods['Net_Search'](cfg, init)
    MixedOp.MODE = None
    torch.manual_seed(888)
    draws = []
    for _ in range(5):
        net.reset_binary_gates()
        draws.append([m.active_index[0] for m in net.redundant_modules])
    ga = torch.Generator().manual_seed(9)
    alphas = []
    for n_, p_ in net.named_alpha_prob_parameters():
        p_.data = p_.data + 0.7 * torch.randn(p_.shape, generator=ga)
        alphas.append(p_.detach().clone())
    torch.manual_seed(889)
    moved = []
    for _ in range(3):
        net.reset_binary_gates()
        moved.append([m.active_index[0] for m in net.redundant_modules])
    rec = {'draws_seed888': np.array(draws), 'draws_seed889_moved': np.array(moved)}
    for i, a in enumerate(alphas):
        rec['alpha%02d' % i] = a
    np.savez_compressed(os.path.join(OUT, 'sampling_seed888.npz'), **npify(rec))


def golden_losses():
    """The ITM loss module of the reference (mmnas/utils/itm_loss.py BCE_Loss, REDUCTION 'sum') and the VGD loss
    expression of train_vgd.py:252-256,320-334, evaluated literally with the torch.nn loss modules the script builds."""
    from mmnas.utils.itm_loss import BCE_Loss
    g = torch.Generator().manual_seed(11)
    pos, negc, negi = (torch.rand(7, generator=g).requires_grad_(True) for _ in range(3))
    loss = BCE_Loss(Bag(REDUCTION='sum'))(pos, negc, negi)
    loss.backward()
    rec = {'itm_pos': pos, 'itm_negc': negc, 'itm_negi': negi, 'itm_loss': loss, 'itm_gpos': pos.grad,
           'itm_gnegc': negc.grad, 'itm_gnegi': negi.grad}
    # --- VGD, the statements of train_vgd.py:252-256 and :320-334 with SCORES_LOSS='kld', LOSS_AVG=True, LOSS_LAMBDA=.5
    b, n = 5, 9
    pred_scores = torch.log_softmax(torch.randn(b, n, generator=g), -1).requires_grad_(True)
    pred_reg = torch.randn(b, n, 4, generator=g).requires_grad_(True)
    train_scores = torch.softmax(torch.randn(b, n, generator=g), -1) * (torch.rand(b, n, generator=g) > 0.5)
    train_scores_mask = torch.tensor([[1.], [1.], [0.], [1.], [1.]])
    train_transformed_bbox = torch.randn(b, n, 4, generator=g)
    train_bbox_mask = (torch.rand(b, n, 1, generator=g) > 0.6).float()
    scores_loss = torch.nn.KLDivLoss(reduction='sum')
    reg_loss = torch.nn.SmoothL1Loss(reduction='sum')
    loss_scores = scores_loss(pred_scores * train_scores_mask, train_scores * train_scores_mask)
    loss_reg = reg_loss(pred_reg * train_bbox_mask, train_transformed_bbox * train_bbox_mask)
    avg_scores = torch.sum(train_scores_mask.data)
    avg_reg = torch.sum(train_bbox_mask.data)
    loss_scores /= avg_scores
    loss_reg /= avg_reg
    loss = loss_scores + 0.5 * loss_reg
    loss.backward()
    rec.update({'vgd_pred_scores': pred_scores, 'vgd_pred_reg': pred_reg, 'vgd_scores': train_scores,
                'vgd_scores_mask': train_scores_mask, 'vgd_tbox': train_transformed_bbox, 'vgd_bbox_mask': train_bbox_mask,
                'vgd_loss': loss, 'vgd_gscores': pred_scores.grad, 'vgd_greg': pred_reg.grad})
    np.savez_compressed(os.path.join(OUT, 'losses.npz'), **npify(rec))


def golden_geometry():
    """relation_embedding (load_data_vqa.py:7-33).  The loader module needs spaCy, so only that
    function's source is extracted from the unmodified file and executed."""
    src = open(os.path.join(REF, 'mmnas/loader/load_data_vqa.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'relation_embedding'][0]
    ns = {'torch': torch}
    exec(compile(ast.Module([fn], []), 'load_data_vqa.py', 'exec'), ns)
    g = torch.Generator().manual_seed(5)
    w, hh = 8 + 192 * torch.rand(9, generator=g), 8 + 152 * torch.rand(9, generator=g)
    x1, y1 = (640 - w) * torch.rand(9, generator=g), (480 - hh) * torch.rand(9, generator=g)
    boxes = torch.stack([x1, y1, x1 + w, y1 + hh], 1)
    boxes[3] = boxes[2]                                    # identical boxes exercise the 1e-3 clamp
    np.savez_compressed(os.path.join(OUT, 'geometry.npz'), boxes=boxes.numpy(),
                        rel=ns['relation_embedding'](boxes).numpy())


def main():
    assert os.path.isdir(REF), 'the reference is only mounted in the authoring container'
    sys.path.insert(0, REF)
    from mmnas.model.mixed import MixedOp
    from mmnas.utils.ops_adapter import OpsAdapter
    from mmnas.model.full_vqa import Net_Full
    from mmnas.model.hygr_vqa import Net_Search
    mods = dict(MixedOp=MixedOp, OpsAdapter=OpsAdapter, Net_Full=Net_Full, Net_Search=Net_Search)
    torch.set_num_threads(1)
    jobs = {'ops': lambda: golden_ops(mods), 'mixed': lambda: golden_mixed(mods), 'net_full': lambda: golden_net_full(mods),
            'net_search': lambda: golden_net_search(mods), 'vgd': lambda: golden_net_full_task('vgd'),
            'itm': lambda: golden_net_full_task('itm'), 'geometry': golden_geometry,
            'mixed_two': lambda: golden_mixed_two(mods), 'sampling': lambda: golden_sampling(mods),
            'losses': golden_losses}
    only = [a for a in sys.argv[1:] if a in jobs] or list(jobs)      # `make_golden.py mixed_two sampling` regenerates two
    for name in only:
        jobs[name]()
    for f in sorted(os.listdir(OUT)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
