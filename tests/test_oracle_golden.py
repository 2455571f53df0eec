"""Pins oracle/mmnas_oracle.py against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py) and, when /root/reference is mounted, against the live reference."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import mmnas_oracle as O
from tests.util import load_golden, params_of, literal, normwise, grad_floor

TOL = 2e-6   # same torch CPU kernels, same op order: only reassociation noise


@pytest.mark.parametrize('name', ['self_att_64', 'rel_self_att_64', 'guided_att_64', 'feed_forward'])
def test_ops_match_golden(name):
    r = load_golden('ops_h128.npz', name)
    P = O.leaf_params(params_of(r))
    x, y, rel = (r[k].clone().requires_grad_(True) for k in ('x', 'y', 'rel'))
    out = O.op_forward(name, P, '', x, y, r['x_mask'], r['y_mask'], rel)
    assert normwise(out, r['out']) < TOL
    out.backward(r['gout'])
    assert normwise(x.grad, r['gx']) < TOL
    if 'gy' in r:
        assert normwise(y.grad, r['gy']) < TOL
    if 'grel' in r:
        assert normwise(rel.grad, r['grel']) < TOL
    for k, p in P.items():
        assert normwise(p.grad, r['g.' + k]) < TOL, k


def test_fully_padded_sample_is_uniform_attention():
    r = load_golden('ops_h128.npz', 'self_att_64')
    assert bool(r['x_mask'][2].all())          # the golden case really contains an all-padded sample
    assert torch.isfinite(r['out']).all()


def test_mixed_full_mode_matches_golden():
    r = load_golden('mixed_h64.npz')
    P = O.leaf_params(params_of(r))
    x, y = (r[k].clone().requires_grad_(True) for k in ('x', 'y'))
    out = O.mixed_forward(O.DEC_SAFE, P, '', x, y, r['x_mask'], r['y_mask'], r['rel'], 'full',
                          r['active'].tolist(), r['inactive'].tolist())
    assert normwise(out, r['out']) < TOL
    out.backward(r['gout'])
    assert normwise(x.grad, r['gx']) < TOL
    if 'gy' in r:
        assert normwise(y.grad, r['gy']) < TOL
    else:
        assert y.grad is None                  # the sampled active candidate does not read `pre`
    assert normwise(P['alpha_gate'].grad, r['gate_grad']) < TOL
    assert normwise(O.arch_param_grad(P['alpha_prob'], P['alpha_gate'].grad), r['prob_grad']) < TOL
    a = int(r['active'][0])
    for k, p in P.items():
        if k.startswith('candidate_ops.%d.' % a):
            assert normwise(p.grad, r['g.' + k]) < TOL, k
        elif k.startswith('candidate_ops.'):
            assert p.grad is None, k           # inactive candidates are detached (mixed.py:66-68)


def test_net_full_step_matches_golden():
    r = load_golden('net_full_h64.npz')
    P = O.leaf_params(params_of(r))
    inputs = (r['frcn'], r['bbox'], r['rel'], r['ques'], r['rel_q'])
    loss, pred = O.train_step_vqa(P, inputs, r['target'], literal(r, 'genotype'))
    assert normwise(pred, r['pred']) < TOL
    assert abs(loss.item() - r['loss'].item()) < 1e-5 * abs(r['loss'].item())
    for k, p in P.items():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert normwise(g, r['g.' + k], grad_floor(r)) < 5e-6, k


def test_net_search_arch_step_matches_golden():
    r = load_golden('net_search_h64.npz')
    P = O.leaf_params(params_of(r))
    inputs = (r['frcn'], r['bbox'], r['rel'], r['ques'], r['rel_q'])
    choices = {'enc': r['choices_enc'].tolist(), 'dec': r['choices_dec'].tolist()}
    pred = O.net_search_vqa(P, inputs, 'full', choices, n_enc=2, n_dec=3)
    assert normwise(pred, r['pred']) < TOL
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, r['target'], reduction='sum')
    loss.backward()
    for k in P:
        if k.endswith('alpha_gate'):
            # <o_i, dOut> is a cancelling sum: fp32 reassociation noise (manual LSTM recurrence here vs
            # torch's fused LSTM in the reference) shows up amplified, measured 6e-6
            assert normwise(P[k].grad, r['g.' + k]) < 5e-5, k
            kp = k.replace('alpha_gate', 'alpha_prob')
            pg = O.arch_param_grad(P[kp], P[k].grad)
            assert normwise(pg, r['g.' + kp]) < 5e-5, kp
            # alpha Adam (lr 0.1, betas (0, .999)): first step moves each alpha by -0.1*sign(grad)
            after = P[kp].detach() - 0.1 * pg / (pg.abs() + 1e-8 * (1 - 0.999) ** 0.5)
            assert normwise(after, r['after.' + kp]) < 1e-5, kp


def test_geometry_matches_golden():
    z = load_golden('geometry.npz')
    assert torch.equal(O.relation_embedding(z['boxes']), z['rel'])


@pytest.mark.skipif(not os.path.isdir('/root/reference/mmnas'), reason='reference not mounted')
def test_oracle_matches_live_reference_at_bench_shapes():
    """Not only the committed vectors: the oracle equals the reference at the real op shapes (B=4 of config T)."""
    sys.path.insert(0, '/root/reference')
    try:
        from mmnas.utils.ops_adapter import OpsAdapter
    finally:
        sys.path.remove('/root/reference')

    class C:
        HSIZE, DROPOUT_R, REL_SIZE = 512, 0.0, 64
    torch.manual_seed(0)
    b, nx, ny = 4, 100, 14
    x, y = torch.randn(b, nx, 512), torch.randn(b, ny, 512)
    rel = torch.relu(torch.randn(b, nx, nx, 64))
    xm = torch.zeros(b, 1, 1, nx, dtype=torch.bool); xm[1, ..., 37:] = True
    ym = torch.zeros(b, 1, 1, ny, dtype=torch.bool); ym[2, ..., 5:] = True
    for name in ('self_att_64', 'rel_self_att_64', 'guided_att_64', 'feed_forward'):
        op = OpsAdapter().OPS[name](C, True, True)
        ref = op(x, y, xm, ym, rel)
        out = O.op_forward(name, dict(op.state_dict()), '', x, y, xm, ym, rel)
        assert normwise(out, ref) < TOL, name


@pytest.mark.parametrize('task', ['vgd', 'itm'])
def test_net_full_vgd_itm_match_golden(task):
    r = load_golden('net_full_%s_h64.npz' % task)
    P = O.leaf_params(params_of(r))
    inputs = (r['frcn'], r['bbox'], r['rel'], r['ques'], r['rel_q'])
    outs = O.net_full(P, inputs, literal(r, 'genotype'), task=task)
    outs = outs if isinstance(outs, tuple) else (outs,)
    for i, o in enumerate(outs):
        assert normwise(o, r['out%d' % i]) < TOL
    sum((o * r['w%d' % i]).sum() for i, o in enumerate(outs)).backward()
    floor = grad_floor(r)
    for k, p in P.items():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert normwise(g, r['g.' + k], floor) < 5e-6, k


def test_mixed_two_mode_matches_golden():
    """MODE 'two' (mixed.py:136-148, :179-191, :200-208): pair sampling, gated output, 2 x 2 gradient rule, rescale."""
    r = load_golden('mixed_two_h64.npz')
    P = O.leaf_params(params_of(r))
    torch.manual_seed(888)
    active, inactive = O.binarize_two(P['alpha_prob'])
    assert [active] == r['active'].tolist() and [inactive] == r['inactive'].tolist()
    x, y = (r[k].clone().requires_grad_(True) for k in ('x', 'y'))
    out = O.mixed_forward(O.DEC_SAFE, P, '', x, y, r['x_mask'], r['y_mask'], r['rel'], 'two', [active], [inactive])
    assert normwise(out, r['out']) < TOL
    out.backward(r['gout'])
    assert normwise(x.grad, r['gx']) < TOL
    assert normwise(P['alpha_gate'].grad, r['gate_grad']) < TOL
    g = O.arch_param_grad_two(P['alpha_prob'], P['alpha_gate'].grad, active, inactive)
    assert normwise(g, r['prob_grad']) < TOL
    assert normwise(O.rescale_two(r['alpha_adam'], r['alpha_before'], active, inactive), r['alpha_rescaled']) < TOL
    untouched = [i for i in range(4) if i not in (active, inactive)]
    assert torch.equal(r['alpha_rescaled'][untouched], r['alpha_before'][untouched])


def test_losses_match_golden():
    """BCE_Loss of mmnas/utils/itm_loss.py (positive term twice) and the VGD loss of train_vgd.py:320-334."""
    r = load_golden('losses.npz')
    pos, negc, negi = (r[k].clone().requires_grad_(True) for k in ('itm_pos', 'itm_negc', 'itm_negi'))
    loss = O.itm_bce_loss(pos, negc, negi)
    loss.backward()
    assert normwise(loss, r['itm_loss']) < TOL
    for t, k in ((pos, 'itm_gpos'), (negc, 'itm_gnegc'), (negi, 'itm_gnegi')):
        assert normwise(t.grad, r[k]) < TOL
    assert normwise(pos.grad, 2 * (-1 / r['itm_pos'])) < TOL          # the double-counted positive term
    ps, pr = (r[k].clone().requires_grad_(True) for k in ('vgd_pred_scores', 'vgd_pred_reg'))
    loss = O.vgd_loss(ps, pr, r['vgd_scores'], r['vgd_scores_mask'], r['vgd_tbox'], r['vgd_bbox_mask'])
    loss.backward()
    assert normwise(loss, r['vgd_loss']) < TOL
    assert normwise(ps.grad, r['vgd_gscores']) < TOL and normwise(pr.grad, r['vgd_greg']) < TOL
