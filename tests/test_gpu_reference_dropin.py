"""The drop-in claim on hardware: the reference's OWN callers — mmnas/model/full_vqa.py Net_Full (genotype from the
reference's arch/mmnas_vqa.json), mmnas/model/hygr_vqa.py Net_Search with its supernet bookkeeping, and the VGD / ITM
nets full_vgd.py / full_itm.py (arch/mmnas_vgd.json, arch/mmnas_itm.json) through the step bodies of train_vgd.py /
train_itm.py — run on this library's CUDA operators through mmnas_b200.install_as_mmnas(), against the untouched
reference (its PyTorch ops) on the same GPU, float32: logits, loss and gradients within the fp32 tolerance of
north_star, same sampled path under the same seed, same genotype.  The reference files are the byte-identical copy
under baseline/_ref/ (git-ignored, populated by scripts/install_reference.py / __graft_entry__.build(); the manifest's
sha256 sums are re-checked)."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

from tests.util import Parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def runs():
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    import install_reference
    install_reference.install(verbose=False)          # no-op on the GPU box (no /root/reference there)
    if not install_reference.verify():
        pytest.fail('baseline/_ref is missing or modified: run scripts/install_reference.py in the authoring container')
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for impl in ('ref', 'ours', 'ours_bf16'):
            path = os.path.join(tmp, impl + '.pt')
            r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', '_dropin_worker.py'), impl, path],
                               capture_output=True, text=True, cwd=ROOT)
            assert r.returncode == 0, r.stderr[-3000:]
            out[impl] = torch.load(path, weights_only=False)
    return out


def _tol(name):
    # fp32-vs-fp32: a ReLU pre-activation within rounding of 0 may take the other branch in the two evaluations
    # (moves one token's contribution to dW1 / db1); linear_r.bias is a cancelling sum (tests/test_gpu_blocks.py)
    return 1e-3 if ('mlp.fc.linear' in name or name.endswith('linear_r.bias')) else 3e-5


def test_reference_net_full_runs_on_the_cuda_operators(runs):
    ref, ours = runs['ref']['full'], runs['ours']['full']
    assert runs['ref']['modules'] == 'mmnas.model.modules' and runs['ours']['modules'] == 'mmnas_b200.model.modules'
    assert runs['ours']['launches'] > 300                      # the C ABI really ran (30 blocks fwd + bwd)
    assert ours['keys'] == ref['keys']                         # checkpoint-compatible state dict
    pr = Parity('dropin/full_vqa.Net_Full/fp32')
    pr.add('pred', ours['pred'], ref['pred'], 1e-5)
    pr.add('loss', ours['loss'], ref['loss'], 1e-5)
    floor = 1e-2 * max(float(g.abs().max()) for g in ref['grads'].values())
    for n_, g in ref['grads'].items():
        pr.add(n_, ours['grads'][n_], g, _tol(n_), floor)
    pr.check()


def test_reference_net_search_arch_step_runs_on_the_cuda_operators(runs):
    ref, ours = runs['ref']['search'], runs['ours']['search']
    assert ours['picks'] == ref['picks']                       # same seed -> same sampled path through OUR MixedOp
    assert ours['genotype'] == ref['genotype']
    pr = Parity('dropin/hygr_vqa.Net_Search/full/fp32')
    pr.add('pred', ours['pred'], ref['pred'], 1e-5)
    pr.add('loss', ours['loss'], ref['loss'], 1e-5)
    gfloor = 1e-2 * max(float(g.abs().max()) for g in ref['gate'].values())
    for n_, g in ref['gate'].items():
        pr.add(n_ + '.grad', ours['gate'][n_], g, 1.5e-4, gfloor)      # <o_k, dOut> over 1.6 M elements: cancelling sums
    for n_, g in ref['prob'].items():
        pr.add(n_ + '.grad', ours['prob'][n_], g, 1.5e-4, gfloor)
    floor = 1e-2 * max(float(g.abs().max()) for g in ref['grads'].values())
    assert set(ours['grads']) == set(ref['grads'])
    for n_, g in ref['grads'].items():
        pr.add(n_, ours['grads'][n_], g, _tol(n_), floor)
    pr.check()


def _train_net_parity(label, ref, ours, extra=()):
    assert ours['keys'] == ref['keys']                         # checkpoint-compatible state dict
    pr = Parity(label)
    pr.add('pred', ours['pred'], ref['pred'], 1e-5)
    for k in extra:
        pr.add(k, ours[k], ref[k], 1e-5)
    pr.add('loss', ours['loss'], ref['loss'], 1e-5)
    floor = 1e-2 * max(float(g.abs().max()) for g in ref['grads'].values())
    assert set(ours['grads']) == set(ref['grads'])
    for n_, g in ref['grads'].items():
        pr.add(n_, ours['grads'][n_], g, _tol(n_), floor)
    pr.check()


def test_reference_vgd_net_runs_on_the_cuda_operators(runs):
    """BASELINE configs[3]'s caller: full_vgd.Net_Full from arch/mmnas_vgd.json (6 RSA blocks, grounding head with
    region scores + box regression), loss of train_vgd.py:320-334."""
    _train_net_parity('dropin/full_vgd.Net_Full/fp32', runs['ref']['vgd'], runs['ours']['vgd'], extra=('pred_reg',))


def test_reference_itm_net_runs_on_the_cuda_operators(runs):
    """BASELINE configs[4]'s caller: full_itm.Net_Full from arch/mmnas_itm.json, three forwards of one net before one
    backward (train_itm.py:387-391) and the reference's BCE_Loss — every operator instance holds three live
    workspaces at once."""
    _train_net_parity('dropin/full_itm.Net_Full/fp32', runs['ref']['itm'], runs['ours']['itm'])


@pytest.mark.parametrize('task', ['vgd', 'itm'])
def test_reference_search_nets_weight_step_runs_on_the_cuda_operators(runs, task):
    """hygr_vgd.Net_Search / hygr_itm.Net_Search (the supernets of search_vgd.py / search_itm.py): the reference's own
    bookkeeping samples a path under seed 888 through OUR MixedOp, switches the unused candidates off, and one weight-step
    body runs on the CUDA operators — same path, same scores / loss, same gradients on every sampled candidate."""
    ref, ours = runs['ref']['search_' + task], runs['ours']['search_' + task]
    assert ours['picks'] == ref['picks']
    pr = Parity('dropin/hygr_%s.Net_Search/weight_step/fp32' % task)
    pr.add('pred', ours['pred'], ref['pred'], 1e-5)
    pr.add('loss', ours['loss'], ref['loss'], 1e-5)
    floor = 1e-2 * max(float(g.abs().max()) for g in ref['grads'].values())
    assert set(ours['grads']) == set(ref['grads'])
    for n_, g in ref['grads'].items():
        pr.add(n_, ours['grads'][n_], g, _tol(n_), floor)
    pr.check()


@pytest.mark.parametrize('task', ['full', 'vgd', 'itm'])
def test_reference_train_nets_run_on_the_bf16_arm(runs, task):
    """The same reference callers with mmnas_b200.set_precision('bf16') (the tcgen05 kernels: what a user switching to
    this library actually trains with) against the untouched float32 reference: logits / scores and loss within
    north_star's 2e-2, gradients within the bf16 gate of tests/test_gpu_nets.py (5e-2 Frobenius-relative)."""
    ref, ours = runs['ref'][task], runs['ours_bf16'][task]
    assert runs['ours_bf16']['modules'] == 'mmnas_b200.model.modules'
    assert ours['keys'] == ref['keys']
    pr = Parity('dropin/%s.Net_Full/bf16' % {'full': 'full_vqa', 'vgd': 'full_vgd', 'itm': 'full_itm'}[task])
    pr.add('pred', ours['pred'], ref['pred'], 2e-2)
    if task == 'vgd':
        pr.add('pred_reg', ours['pred_reg'], ref['pred_reg'], 2e-2)
    pr.add('loss', ours['loss'], ref['loss'], 2e-2)
    floor = 1e-2 * max(float(g.abs().max()) for g in ref['grads'].values())
    assert set(ours['grads']) == set(ref['grads'])
    for n_, g in ref['grads'].items():
        pr.add(n_, ours['grads'][n_], g, 5e-2, floor, metric='fro')
    pr.check()
