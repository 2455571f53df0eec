"""Host-side logic of the product path that needs no GPU: architecture sampling with the reference's random-number
consumption, the MODE 'two' bookkeeping, the task losses and the synthetic batch shapes — against golden vectors made
from the unmodified reference (tests/golden/make_golden.py: sampling, mixed_two, losses)."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, params_of, normwise


def tiny_search_net():
    from mmnas_b200.data.synthetic import Cfg
    from mmnas_b200.model.nets import Net_Search
    cfg = Cfg(mode='search', HSIZE=64, FRCNFEAT_SIZE=32, BBOXFEAT_EMB_SIZE=32, WORD_EMBED_SIZE=16, ATTFLAT_MLP_SIZE=48,
              ATTFLAT_OUT_SIZE=128)
    torch.manual_seed(888)
    return Net_Search(cfg, {'token_size': 30, 'ans_size': 11, 'pretrained_emb': np.zeros((30, 16), np.float32)})


@pytest.mark.parametrize('batched', [False, True])
def test_seed_888_draws_the_reference_architecture_samples(batched):
    """search_vqa.py:62 seeds 888; Net_Search.reset_binary_gates (hygr_vqa.py:168-172) then draws one multinomial per
    MixedOp.  Both the per-module path and the step harness's batched path must reproduce the reference's samples —
    five consecutive draws, and three more after the alphas moved."""
    r = load_golden('sampling_seed888.npz')
    net = tiny_search_net()
    torch.manual_seed(888)
    draws = []
    for _ in range(5):
        net.reset_binary_gates(batched=batched)
        draws.append([m.active_index[0] for m in net.redundant_modules])
        for m in net.redundant_modules:
            a = m.active_index[0]
            assert m.alpha_gate.data[a] == 1 and m.alpha_gate.data.sum() == 1
            assert sorted(m.active_index + m.inactive_index) == list(range(m.n_choices))
    assert draws == r['draws_seed888'].tolist()
    for i, (_, p) in enumerate(net.named_alpha_prob_parameters()):
        p.data = r['alpha%02d' % i].clone()
    torch.manual_seed(889)
    moved = []
    for _ in range(3):
        net.reset_binary_gates(batched=batched)
        moved.append([m.active_index[0] for m in net.redundant_modules])
    assert moved == r['draws_seed889_moved'].tolist()


def test_mixed_op_two_mode_bookkeeping_matches_reference():
    """binarize() in MODE 'two' draws the reference's pair under the same seed; set_arch_param_grad applies the 2 x 2
    rule to the pair only; rescale_updated_arch_param restores the pair's logsumexp after the alpha step."""
    from mmnas_b200.model.mixed import MixedOp

    class C:
        HSIZE, DROPOUT_R, REL_SIZE, OPS_NORM, OPS_RESIDUAL = 64, 0.0, 64, True, True
    r = load_golden('mixed_two_h64.npz')
    m = MixedOp(C, 'dec_safe')
    m.load_state_dict(params_of(r))
    opt = torch.optim.Adam([m.alpha_prob], 0.1, betas=(0., 0.999), weight_decay=0)
    MixedOp.MODE = 'two'
    try:
        torch.manual_seed(888)
        m.binarize()
        assert m.active_index == r['active'].tolist() and m.inactive_index == r['inactive'].tolist()
        assert m.alpha_gate.data[m.active_index[0]] == 1 and m.alpha_gate.data.sum() == 1
        m.alpha_gate.grad = r['gate_grad'].clone()
        m.set_arch_param_grad()
        assert normwise(m.alpha_prob.grad, r['prob_grad']) < 2e-6
        opt.step()
        assert normwise(m.alpha_prob.data, r['alpha_adam']) < 2e-6
        m.rescale_updated_arch_param()
        assert normwise(m.alpha_prob.data, r['alpha_rescaled']) < 2e-6
    finally:
        MixedOp.MODE = None


def test_task_losses_match_reference():
    from mmnas_b200.engine import itm_loss, vgd_loss
    r = load_golden('losses.npz')
    pred = torch.cat([r['itm_pos'], r['itm_negc'], r['itm_negi']]).requires_grad_(True)
    loss = itm_loss(pred)
    loss.backward()
    assert normwise(loss, r['itm_loss']) < 2e-6
    assert normwise(pred.grad, torch.cat([r['itm_gpos'], r['itm_gnegc'], r['itm_gnegi']])) < 2e-6
    ps, pr = (r[k].clone().requires_grad_(True) for k in ('vgd_pred_scores', 'vgd_pred_reg'))
    loss = vgd_loss((ps, pr), (r['vgd_scores'], r['vgd_scores_mask'], r['vgd_tbox'], r['vgd_bbox_mask']))
    loss.backward()
    assert normwise(loss, r['vgd_loss']) < 2e-6
    assert normwise(ps.grad, r['vgd_gscores']) < 2e-6 and normwise(pr.grad, r['vgd_greg']) < 2e-6


def test_synthetic_batches_have_the_baseline_shapes():
    from mmnas_b200.data.synthetic import make_batch, spec_for
    (f, b, rel, q, rq), t = make_batch(spec_for('vqa', batch=3))
    assert f.shape == (3, 100, 2048) and rel.shape == (3, 100, 100, 4) and q.shape == (3, 14) and t.shape == (3, 3129)
    (f, b, rel, q, rq), t = make_batch(spec_for('vgd', batch=3))
    assert q.shape == (3, 15) and bool((q[:, -1] == 0).all())            # load_data_vgd.py:190: max_token + 1 slots
    assert bool((f.abs().sum(-1) > 0).all())                              # every region valid
    assert [x.shape for x in t] == [(3, 100), (3, 1), (3, 100, 4), (3, 100, 1)]
    (f, b, rel, q, rq), t = make_batch(spec_for('itm', batch=3))
    assert f.shape == (9, 36, 2048) and q.shape == (9, 50)
    assert torch.equal(f[0:3], f[3:6]) and torch.equal(q[0:3], q[6:9])   # negc keeps the images, negi the captions
    assert not torch.equal(q[0:3], q[3:6]) and not torch.equal(f[0:3], f[6:9])


def test_sink_prepare_is_refused_inside_a_side_stream_fork():
    """functional._Fork moves only the library's launches to the side stream; a torch fill issued inside the block would
    run on the main stream and race them (the LSTM weight-gradient bug of round 2).  _Sink.prepare() must refuse."""
    import pytest
    from mmnas_b200 import functional as Fn
    w = torch.nn.Parameter(torch.zeros(4, 3))
    sink = Fn._Sink((w,), torch.device('cpu'))
    sink.prepare(zero=True)                       # outside a fork: fine
    assert sink.buf.shape == (4, 3) and not sink.direct
    Fn._Fork.depth += 1
    try:
        with pytest.raises(RuntimeError, match='inside'):
            Fn._Sink((w,), torch.device('cpu')).prepare(zero=True)
    finally:
        Fn._Fork.depth -= 1
