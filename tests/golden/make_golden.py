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
    golden_ops(mods)
    golden_mixed(mods)
    golden_net_full(mods)
    golden_net_search(mods)
    golden_net_full_task('vgd')
    golden_net_full_task('itm')
    golden_geometry()
    for f in sorted(os.listdir(OUT)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
