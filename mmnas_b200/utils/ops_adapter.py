"""Operator registry — the plug-in boundary of the hot path (drop-in for mmnas/utils/ops_adapter.py:5-74).

`OpsAdapter().OPS[name](__C, norm, residual)` builds a candidate block; `OpsAdapter().Used_OPS[kind]` lists the
candidates of an encoder / decoder supernet node.  Every name of the reference registry that maps onto the
four block families implemented in CUDA (self_att_*, rel_self_att_*, guided_att_*, feed_forward*) is
registered, plus 'none' / 'skip_connect'.  The remaining reference names (conv, GLU, uni-image attention,
activation-only ops) appear in no Used_OPS list and no arch/*.json; asking for one raises KeyError with
an explanation instead of silently building something else.
"""
from ..model import modules as M

ENC_CANDIDATES = ['self_att_64', 'feed_forward']
DEC_CANDIDATES = ['self_att_64', 'rel_self_att_64', 'guided_att_64', 'feed_forward']

_NOT_BUILT = ('relu', 'gelu', 'leakyrelu', 'uniimg_att_128', 'uniimg_att_64', 'uniimg_att_32',
              'sep_conv_3', 'sep_conv_5', 'sep_conv_7', 'sep_conv_11', 'std_conv_3', 'std_conv_5', 'std_conv_7',
              'std_conv_11', 'gated_linear_1', 'gated_linear_2', 'feed_forward_deep')


class _Registry(dict):
    def __missing__(self, key):
        if key in _NOT_BUILT:
            raise KeyError("operator %r exists in the reference registry but is in no search space or arch JSON; "
                           "it is outside the CUDA hot path and not built" % key)
        raise KeyError(key)


def _att(cls, base, hsize_k=None):
    return lambda __C, norm, residual: cls(__C, norm, residual, base=base, hsize_k=hsize_k)


def _ffn(mid_k=None):
    return lambda __C, norm, residual: M.FeedForward(__C, norm, residual, mid_k=mid_k)


class OpsAdapter:
    def __init__(self):
        self.Used_OPS = {'enc_safe': list(ENC_CANDIDATES), 'dec_safe': list(DEC_CANDIDATES)}
        self.Used_OPS['enc'] = self.Used_OPS['enc_safe'] + ['none']
        self.Used_OPS['dec'] = self.Used_OPS['dec_safe'] + ['none']

        ops = _Registry()
        ops['none'] = lambda __C, norm, residual: M.Zero()
        ops['skip_connect'] = lambda __C, norm, residual: M.Identity()
        for base in (256, 128, 64, 32, 16):
            ops['self_att_%d' % base] = _att(M.SelfAtt, base)
            ops['rel_self_att_%d' % base] = _att(M.RelSelfAtt, base)
            ops['guided_att_%d' % base] = _att(M.GuidedAtt, base)
        ops['self_att_64_2'] = _att(M.SelfAtt, 64, hsize_k=2)
        ops['guided_att_64_2'] = _att(M.GuidedAtt, 64, hsize_k=2)
        ops['feed_forward'] = _ffn()
        for k in (2, 8, 16, 32):
            ops['feed_forward_%d' % k] = _ffn(k)
        self.OPS = ops
