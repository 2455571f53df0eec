"""Worker of tests/test_gpu_reference_dropin.py (run as a subprocess, one per implementation).

    python tests/_dropin_worker.py {ours|ours_bf16|ref} OUT.pt

Imports the UNMODIFIED reference callers from baseline/_ref (mmnas/model/full_vqa.py, hygr_vqa.py — populated by
scripts/install_reference.py) and runs, on cuda:0 in float32:
  1. full_vqa.Net_Full with the genotype read from baseline/_ref/arch/mmnas_vqa.json exactly as train_vqa.py:185 does,
     one train-step body (train_vqa.py:294-299): logits, loss, every parameter gradient;
  2. hygr_vqa.Net_Search, one architecture-step body in MODE 'full' (search_vqa.py:317-331) driven through the
     reference's own bookkeeping (reset_binary_gates / unused_modules_off / set_arch_param_grad / genotype);
  3. full_vgd.Net_Full (arch/mmnas_vgd.json, 100 regions x 15 query tokens), the step body of train_vgd.py:317-336
     (KLD over the masked region scores + 0.5 SmoothL1 over the masked box regressions, LOSS_AVG);
  4. full_itm.Net_Full (arch/mmnas_itm.json, 36 regions x 50 caption tokens), the step body of train_itm.py:384-392:
     three forwards (positive, negative captions, negative images) and the reference's own BCE_Loss
     (mmnas/utils/itm_loss.py, positive term twice);
  5. hygr_vgd.Net_Search and hygr_itm.Net_Search, one weight-step body each on the path sampled under seed 888
     (search_vgd.py / search_itm.py: reset_binary_gates, unused_modules_off, forward, loss, backward).
With `ours`, mmnas_b200.install_as_mmnas() first replaces mmnas.model.modules, mmnas.model.mixed and
mmnas.utils.ops_adapter, so the reference's nets run on this library's CUDA operators; with `ref` nothing is replaced.
`ours_bf16` is `ours` in the bf16 arm (tcgen05 kernels; train-time nets 1, 3, 4 only)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')


def main():
    impl, out_path = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ROOT)
    sys.path.insert(1, REF)
    torch.backends.cudnn.allow_tf32 = False           # the callers' cuDNN LSTM in true fp32 for both implementations
    torch.backends.cuda.matmul.allow_tf32 = False
    bf16 = impl == 'ours_bf16'
    if bf16:
        impl = 'ours'
    if impl == 'ours':
        import mmnas_b200
        mmnas_b200.install_as_mmnas()
        mmnas_b200.set_precision('bf16' if bf16 else 'fp32')
    from mmnas.model.full_vqa import Net_Full
    from mmnas.model.hygr_vqa import Net_Search
    from mmnas.model.mixed import MixedOp
    import mmnas.model.modules as M
    from mmnas_b200.data.synthetic import Cfg, SynthSpec, make_batch, init_dict
    from tests.util import condition_rsa_
    dev = os.environ.get('MMNAS_DROPIN_DEV', 'cuda:0')      # ('cpu' only to smoke-test the reference leg without a GPU)
    res = {'modules': M.__name__}

    # ---- 1. train-time net from the reference's arch JSON
    geno = json.load(open(os.path.join(REF, 'arch', 'mmnas_vqa.json')))['epoch0']      # train_vqa.py:185
    spec = SynthSpec(batch=4, vocab=1000, n_ans=100)
    cfg = Cfg(genotype=geno, DROPOUT_R=0.0)
    inputs, target = make_batch(spec, 888)
    torch.manual_seed(888)
    net = Net_Full(cfg, init_dict(spec))
    with torch.no_grad():
        condition_rsa_(dict(net.named_parameters()))
    net = net.to(dev).train()
    pred = net(tuple(t.to(dev) for t in inputs))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, target.to(dev), reduction='sum')
    loss = loss + 0 * sum(p.sum() for p in net.parameters())                           # train_vqa.py:298
    loss.backward()
    res['full'] = {'pred': pred.detach().cpu(), 'loss': loss.detach().cpu(),
                   'grads': {n: p.grad.detach().cpu() for n, p in net.named_parameters()},
                   'keys': {k: tuple(v.shape) for k, v in net.state_dict().items()}}
    del net, pred, loss

    if not bf16:
        search_arch_step(res, Net_Search, MixedOp, Cfg, init_dict, condition_rsa_, spec, inputs, target, dev, impl)
    train_nets(res, Cfg, make_batch, init_dict, condition_rsa_, dev, impl)
    if not bf16:
        search_weight_steps(res, Cfg, make_batch, init_dict, condition_rsa_, dev)
    torch.save(res, out_path)


def search_arch_step(res, Net_Search, MixedOp, Cfg, init_dict, condition_rsa_, spec, inputs, target, dev, impl):
    # ---- 2. supernet, MODE 'full' architecture step through the reference's own Net_Search methods
    cfg = Cfg(mode='search', DROPOUT_R=0.0)
    torch.manual_seed(888)
    net = Net_Search(cfg, init_dict(spec))
    with torch.no_grad():
        condition_rsa_(dict(net.named_parameters()))
    net = net.to(dev).train()
    MixedOp.MODE = 'full'
    torch.manual_seed(888)
    net.reset_binary_gates()
    net.unused_modules_off()
    picks = [m.active_index[0] for m in net.redundant_modules]
    pred = net(tuple(t.to(dev) for t in inputs))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, target.to(dev), reduction='sum')
    loss = loss + 0 * sum(p.sum() for p in net.alpha_prob_parameters())                # search_vqa.py:321-323
    loss = loss + 0 * sum(p.sum() for p in net.alpha_gate_parameters())
    loss = loss + 0 * sum(p.sum() for p in net.net_parameters())
    net.zero_grad()
    loss.backward()
    gate = {n: p.grad.detach().cpu().clone() for n, p in net.named_alpha_gate_parameters()}
    net.set_arch_param_grad()
    prob = {n: p.grad.detach().cpu().clone() for n, p in net.named_alpha_prob_parameters()}
    net.unused_modules_back()
    MixedOp.MODE = None
    res['search'] = {'picks': picks, 'pred': pred.detach().cpu(), 'loss': loss.detach().cpu(), 'gate': gate, 'prob': prob,
                     'genotype': net.genotype(),
                     'grads': {n: p.grad.detach().cpu() for n, p in net.named_net_parameters() if p.grad is not None}}
    if impl == 'ours':
        from mmnas_b200 import _lib
        res['launches'] = _lib.launches()               # kernels of the library launched by 1. and 2.


def train_nets(res, Cfg, make_batch, init_dict, condition_rsa_, dev, impl):
    # ---- 3. VGD train-time net (RSA-heavy decoder, grounding head) through the step body of train_vgd.py:317-336
    from mmnas.model import full_vgd, full_itm
    from mmnas.utils.itm_loss import BCE_Loss
    from mmnas_b200.data.synthetic import spec_for
    geno = json.load(open(os.path.join(REF, 'arch', 'mmnas_vgd.json')))['epoch0']
    spec = spec_for('vgd', batch=4, vocab=1000)
    cfg = Cfg(genotype=geno, DROPOUT_R=0.0, SCORES_LOSS='kld', LOSS_AVG=True, LOSS_LAMBDA=0.5, REDUCTION='sum')
    inputs, target = make_batch(spec, 888)
    torch.manual_seed(888)
    net = full_vgd.Net_Full(cfg, init_dict(spec))
    with torch.no_grad():
        condition_rsa_(dict(net.named_parameters()))
    net = net.to(dev).train()
    scores, scores_mask, tbox, bbox_mask = (t.to(dev) for t in target)
    pred_scores, pred_reg = net(tuple(t.to(dev) for t in inputs))
    loss_scores = torch.nn.KLDivLoss(reduction=cfg.REDUCTION)(pred_scores * scores_mask, scores * scores_mask)
    loss_reg = torch.nn.SmoothL1Loss(reduction=cfg.REDUCTION)(pred_reg * bbox_mask, tbox * bbox_mask)
    loss = loss_scores / torch.sum(scores_mask.data) + cfg.LOSS_LAMBDA * (loss_reg / torch.sum(bbox_mask.data))
    loss.backward()
    res['vgd'] = {'pred': pred_scores.detach().cpu(), 'pred_reg': pred_reg.detach().cpu(), 'loss': loss.detach().cpu(),
                  'grads': {n: p.grad.detach().cpu() for n, p in net.named_parameters() if p.grad is not None},
                  'keys': {k: tuple(v.shape) for k, v in net.state_dict().items()}}
    del net, pred_scores, pred_reg, loss

    # ---- 4. ITM train-time net, three forwards + BCE_Loss (train_itm.py:384-392)
    geno = json.load(open(os.path.join(REF, 'arch', 'mmnas_itm.json')))['epoch0']
    spec = spec_for('itm', batch=2, vocab=1000)
    cfg = Cfg(genotype=geno, DROPOUT_R=0.0, REDUCTION='sum')
    # rows [0,B) positive | [B,2B) negative captions | [2B,3B) negative images.  Batch seed 892: with 888 / 889 one FFN /
    # AttFlat hidden unit has a pre-activation within float32 rounding of zero, so even the reference's own float32 and
    # float64 evaluations disagree on that ReLU branch (2.7e-3 / 1.1e-2 on mlp.fc.linear.weight); with 892 they agree
    # to 2e-6 on every tensor, so the comparison measures the operators and not the coin flip
    inputs, _ = make_batch(spec, 892)
    torch.manual_seed(888)
    net = full_itm.Net_Full(cfg, init_dict(spec))
    with torch.no_grad():
        condition_rsa_(dict(net.named_parameters()))
    net = net.to(dev).train()
    B = spec.batch
    parts = [tuple(t[i * B:(i + 1) * B].to(dev) for t in inputs) for i in range(3)]
    scores_pos, scores_negc, scores_negi = net(parts[0]), net(parts[1]), net(parts[2])
    loss = BCE_Loss(cfg)(scores_pos, scores_negc, scores_negi)
    loss.backward()
    res['itm'] = {'pred': torch.cat((scores_pos, scores_negc, scores_negi)).detach().cpu(), 'loss': loss.detach().cpu(),
                  'grads': {n: p.grad.detach().cpu() for n, p in net.named_parameters() if p.grad is not None},
                  'keys': {k: tuple(v.shape) for k, v in net.state_dict().items()}}
    if impl == 'ours':
        from mmnas_b200 import _lib
        res['launches_total'] = _lib.launches()


def vgd_step_loss(cfg, pred, target):
    """train_vgd.py:319-334 with the shipped settings (SCORES_LOSS 'kld', LOSS_AVG, LOSS_LAMBDA 0.5, REDUCTION 'sum')."""
    pred_scores, pred_reg = pred
    scores, scores_mask, tbox, bbox_mask = target
    loss_scores = torch.nn.KLDivLoss(reduction=cfg.REDUCTION)(pred_scores * scores_mask, scores * scores_mask)
    loss_reg = torch.nn.SmoothL1Loss(reduction=cfg.REDUCTION)(pred_reg * bbox_mask, tbox * bbox_mask)
    return loss_scores / torch.sum(scores_mask.data) + cfg.LOSS_LAMBDA * (loss_reg / torch.sum(bbox_mask.data))


def search_weight_steps(res, Cfg, make_batch, init_dict, condition_rsa_, dev, dtype=torch.float32, seeds=(888, 892)):
    # ---- 5. the VGD / ITM supernets (H = 256, 4 heads), weight step on the path sampled under seed 888
    from mmnas.model import hygr_vgd, hygr_itm
    from mmnas.model.mixed import MixedOp
    from mmnas.utils.itm_loss import BCE_Loss
    from mmnas_b200.data.synthetic import spec_for
    cast = lambda t: t.to(dev, dtype) if t.is_floating_point() else t.to(dev)        # noqa: E731
    for task, mod, seed in (('vgd', hygr_vgd, seeds[0]), ('itm', hygr_itm, seeds[1])):
        spec = spec_for(task, batch=4 if task == 'vgd' else 2, vocab=1000)
        cfg = Cfg(mode='search', DROPOUT_R=0.0, SCORES_LOSS='kld', LOSS_AVG=True, LOSS_LAMBDA=0.5, REDUCTION='sum')
        inputs, target = make_batch(spec, seed)
        torch.manual_seed(888)
        net = mod.Net_Search(cfg, init_dict(spec))
        with torch.no_grad():
            condition_rsa_(dict(net.named_parameters()))
        net = net.to(dev, dtype).train()
        MixedOp.MODE = None
        torch.manual_seed(888)
        net.reset_binary_gates()
        net.unused_modules_off()
        picks = [m.active_index[0] for m in net.redundant_modules]
        if task == 'vgd':
            pred = net(tuple(cast(t) for t in inputs))
            loss = vgd_step_loss(cfg, pred, tuple(cast(t) for t in target))
            pred = torch.cat((pred[0].reshape(-1), pred[1].reshape(-1)))
        else:
            B = spec.batch
            parts = [tuple(cast(t[i * B:(i + 1) * B]) for t in inputs) for i in range(3)]
            scores = [net(p) for p in parts]
            loss = BCE_Loss(cfg)(*scores)
            pred = torch.cat(scores)
        net.zero_grad()
        loss.backward()
        net.unused_modules_back()
        res['search_' + task] = {'picks': picks, 'pred': pred.detach().cpu(), 'loss': loss.detach().cpu(),
                                 'grads': {n: p.grad.detach().cpu() for n, p in net.named_net_parameters()
                                           if p.grad is not None}}
        del net, pred, loss


if __name__ == '__main__':
    main()
