"""CPU-side checks of the boundary: the shared library builds, loads and exports every symbol include/mmnas_b200.h
declares (no compute without a GPU), the ctypes table matches the header's arity, the drop-in classes keep the
reference's interface, and the product path fails loudly instead of falling back."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'mmnas_b200.h')


def declared():
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    out = {}
    for m in re.finditer(r'\b(?:int|const char\*|unsigned long long)\s+(mmnas_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ('', 'void') else len(args.split(','))
    return out


def test_library_builds_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from mmnas_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared()
    assert len(decl) == 36            # 22 primitive entry points + 10 block-level ones + 4 queries (ABI v8)
    for name in decl:
        assert hasattr(lib, name), name
    lib.mmnas_abi_version.restype = ctypes.c_int
    assert lib.mmnas_abi_version() == _lib.ABI_VERSION
    for name, argtypes in _lib.SIGNATURES.items():
        assert decl[name] == len(argtypes), name          # binding arity == header arity
    assert set(_lib.SIGNATURES) | {'mmnas_abi_version', 'mmnas_last_error', 'mmnas_launch_count', 'mmnas_att_block_sizeof',
                                   'mmnas_ffn_block_sizeof'} == set(decl)
    # the ctypes mirrors of the block descriptors have the size the library was compiled with
    lib.mmnas_att_block_sizeof.restype = lib.mmnas_ffn_block_sizeof.restype = ctypes.c_int
    assert lib.mmnas_att_block_sizeof() == ctypes.sizeof(_lib.AttBlock)
    assert lib.mmnas_ffn_block_sizeof() == ctypes.sizeof(_lib.FfnBlock)
    # ... and the same field names, in the header's order
    hdr = open(HEADER).read()
    for cname, cls in (('mmnas_att_block', _lib.AttBlock), ('mmnas_ffn_block', _lib.FfnBlock)):
        body = re.sub(r'/\*.*?\*/', '', hdr[hdr.index('typedef struct %s {' % cname):hdr.index('} %s;' % cname)], flags=re.S)
        names = []
        for stmt in body.split('{', 1)[1].split(';'):
            stmt = stmt.strip()
            if not stmt:
                continue
            decls = stmt.split(',')
            first = decls[0].split()[-1]
            names += [first.lstrip('*')] + [x.strip().lstrip('*') for x in decls[1:]]
        assert names == [f[0] for f in cls._fields_], cname


def test_block_descriptor_errors_are_reported_without_a_gpu():
    from mmnas_b200 import _lib
    d = _lib.AttBlock()
    d.precision, d.B, d.Nq, d.Nk, d.H, d.I, d.R = 1, 2, 10, 10, 128, 100, 0
    with pytest.raises(_lib.MMnasLibraryError, match='multiple of the head dim'):
        _lib.call_block('mmnas_mha_ln_fwd', d)
    d.I, d.R = 128, 64
    with pytest.raises(_lib.MMnasLibraryError, match='relation inputs given'):
        _lib.call_block('mmnas_mha_ln_fwd', d)
    with pytest.raises(_lib.MMnasLibraryError, match='exactly one of rel / g4'):
        _lib.call_block('mmnas_rel_mha_ln_fwd', d)
    d.R = 0
    fwd, bwd = _lib.workspace_bytes(d)
    assert fwd % 256 == 0 and bwd % 256 == 0 and fwd >= 2 * 20 * (3 * 128 + 128) + 4 * 20 * 128
    f = _lib.FfnBlock()
    f.precision, f.M, f.H, f.F = 0, 20, 128, 512
    with pytest.raises(_lib.MMnasLibraryError, match='null input'):
        _lib.call_block('mmnas_ffn_ln_fwd', f)


def test_argument_errors_are_reported_without_a_gpu():
    from mmnas_b200 import _lib
    with pytest.raises(_lib.MMnasLibraryError, match='head dim 64'):
        _lib.call('mmnas_attn_fwd', 0, 1, 1, 4, 4, 32, None, 0, None, 0, None, 0, None, None, None, 0, 1.0, None, 0, 0.0, None)
    with pytest.raises(_lib.MMnasLibraryError, match='multiple of 32'):
        _lib.call('mmnas_gemm_bf16', 8, 30, 8, 1, 8, 0, 1, 8, 0, 1, 30, 0, None, 0, 0, None, 0, 1.0, 1, None, 0, 0.0, None)


class Cfg:
    HSIZE, DROPOUT_R, REL_SIZE, OPS_NORM, OPS_RESIDUAL = 128, 0.1, 64, True, True


def test_no_cpu_fallback():
    from mmnas_b200._lib import MMnasLibraryError
    from mmnas_b200.utils.ops_adapter import OpsAdapter
    op = OpsAdapter().OPS['feed_forward'](Cfg, True, True)
    with pytest.raises(MMnasLibraryError, match='no CPU path'):
        op(torch.randn(2, 5, 128))
    with pytest.raises(KeyError, match='outside the CUDA hot path'):
        OpsAdapter().OPS['sep_conv_3']
    sa32 = OpsAdapter().OPS['self_att_32'](Cfg, True, True)
    with pytest.raises((NotImplementedError, MMnasLibraryError)):
        sa32(torch.randn(2, 5, 128))


def test_registry_and_mixed_op_interface():
    from mmnas_b200.model.mixed import MixedOp
    from mmnas_b200.utils.ops_adapter import OpsAdapter
    A = OpsAdapter()
    assert A.Used_OPS['enc_safe'] == ['self_att_64', 'feed_forward']
    assert A.Used_OPS['dec_safe'] == ['self_att_64', 'rel_self_att_64', 'guided_att_64', 'feed_forward']
    assert A.Used_OPS['dec'][-1] == 'none'
    m = MixedOp(Cfg, 'dec_safe')
    assert str(m).startswith('MixedOp') and m.n_choices == 4
    torch.manual_seed(888)
    m.binarize()
    assert len(m.active_index) == 1 and sorted(m.active_index + m.inactive_index) == [0, 1, 2, 3]
    assert float(m.alpha_gate.data.sum()) == 1.0 and float(m.alpha_gate.data[m.active_index[0]]) == 1.0
    m.alpha_gate.grad = torch.tensor([0.3, -0.2, 0.1, 0.05])
    m.set_arch_param_grad()
    p = torch.softmax(m.alpha_prob.data, 0)
    g = m.alpha_gate.grad
    expect = torch.stack([sum(g[j] * p[j] * ((1 if i == j else 0) - p[i]) for j in range(4)) for i in range(4)])
    assert torch.allclose(m.alpha_prob.grad, expect, atol=1e-7)
    saved = m.candidate_ops[2]
    m.candidate_ops[2] = None                       # Net_Search.unused_modules_off / _back protocol
    m.candidate_ops[2] = saved
    m.set_chosen_op_active()
    assert m.active_index == [m.chosen_index[0]]


REFERENCE_CALLERS = [('full_vqa', 'Net_Full'), ('full_vgd', 'Net_Full'), ('full_itm', 'Net_Full'),
                     ('hygr_vqa', 'Net_Search'), ('hygr_vgd', 'Net_Search'), ('hygr_itm', 'Net_Search')]


@pytest.fixture(scope='module')
def reference_callers_built():
    """Builds the reference's six callers twice in subprocesses: untouched ('ref') and after install_as_mmnas() ('ours')."""
    import json
    import subprocess
    code = r'''
import sys, json, importlib, numpy as np, torch
sys.path.insert(0, %r)
mode = sys.argv[1]
sys.path.insert(1, "/root/reference")
if mode == "ours":
    import mmnas_b200; mmnas_b200.install_as_mmnas()
import mmnas.model.modules as M
class C: pass
c = C()
c.__dict__.update(HSIZE=128, DROPOUT_R=0.1, REL_SIZE=64, OPS_NORM=True, OPS_RESIDUAL=True, LAYERS=1, BBOX_FEATURE=False,
    FRCNFEAT_SIZE=32, BBOXFEAT_EMB_SIZE=32, WORD_EMBED_SIZE=16, ATTFLAT_GLIMPSES=1, ATTFLAT_OUT_SIZE=256, ATTFLAT_MLP_SIZE=48,
    SCORES_LOSS="kld", ALPHA_INIT_TYPE="normal", NODES={"enc": 2, "dec": 3},
    GENOTYPE={"enc": [["self_att_64"], ["feed_forward"]], "dec": [["guided_att_64"], ["rel_self_att_64"], ["feed_forward"]]})
out = {"module": M.__name__}
for net_module, net_class in %r:
    Net = getattr(importlib.import_module("mmnas.model." + net_module), net_class)
    torch.manual_seed(888)
    net = Net(c, {"token_size": 20, "ans_size": 7, "pretrained_emb": np.zeros((20, 16), np.float32)})
    sd = net.state_dict()
    out[net_module] = {"keys": {k: list(v.shape) for k, v in sd.items()},
                       "sum": float(sum(v.double().abs().sum() for v in sd.values()))}
print(json.dumps(out))
''' % (ROOT, REFERENCE_CALLERS)
    outs = {}
    for mode in ('ours', 'ref'):
        r = subprocess.run([sys.executable, '-c', code, mode], capture_output=True, text=True, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    return outs


@pytest.mark.skipif(not os.path.isdir('/root/reference/mmnas'), reason='reference not mounted')
@pytest.mark.parametrize('net_module', [m for m, _ in REFERENCE_CALLERS])
def test_drop_in_under_the_reference_nets(reference_callers_built, net_module):
    """install_as_mmnas(): each of the reference's own six callers (full_* train nets, hygr_* supernets) builds on the
    CUDA-backed operators, and its state-dict keys / shapes / same-seed initial values equal those of the untouched
    reference (checkpoint compatibility, train_vqa.py:249).  Construction only — the forward needs a GPU
    (tests/test_gpu_reference_dropin.py)."""
    outs = reference_callers_built
    assert outs['ours']['module'] == 'mmnas_b200.model.modules' and outs['ref']['module'] == 'mmnas.model.modules'
    ours, ref = outs['ours'][net_module], outs['ref'][net_module]
    assert ours['keys'] == ref['keys']
    assert abs(ours['sum'] - ref['sum']) < 1e-6 * ref['sum']   # same seed -> same init
