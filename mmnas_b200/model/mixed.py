"""Supernet node: K candidate blocks weighted by architecture parameters (drop-in for mmnas/model/mixed.py).

Interface kept from the reference (mixed.py:36-208) because Net_Search drives it by name: class attribute
`MODE`, `candidate_ops` (ModuleList whose entries may be set to None and restored), parameters `alpha_prob`
and `alpha_gate`, `active_index` / `inactive_index`, `n_choices`, `Used_OPS`, `probs_over_ops`, `active_op`,
`chosen_index`, `binarize()`, `set_chosen_op_active()`, `set_arch_param_grad()`,
`rescale_updated_arch_param()`; `repr()` starts with 'MixedOp' (hygr_vqa.py:163).

MODE None   -> only the sampled candidate runs (weight step, mixed.py:103-104)
MODE 'full' -> every candidate runs; out = sum_k gate_k * o_k with inactive candidates detached
               (mixed.py:60-68).  The weighted sum and, in backward, all K dot products <o_k, dOut>
               (= alpha_gate.grad) plus the active candidate's gradient come from ONE fused kernel pass each
               (mmnas_mixed_accum / mmnas_mixed_alpha_dot).
The K-element alpha bookkeeping (softmax, sampling, the arch-gradient rule) stays on tiny torch tensors.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..functional import MixedSumFn
from ..utils.ops_adapter import OpsAdapter

OPS_ADAPTER = OpsAdapter()


class MixedOp(nn.Module):
    MODE = None

    def __init__(self, __C, name):
        super().__init__()
        self.Used_OPS = OPS_ADAPTER.Used_OPS.get(name, [name])
        self.n_choices = len(self.Used_OPS)
        self.candidate_ops = nn.ModuleList(
            OPS_ADAPTER.OPS[op](__C, norm=__C.OPS_NORM, residual=__C.OPS_RESIDUAL) for op in self.Used_OPS)
        self.alpha_prob = nn.Parameter(torch.zeros(self.n_choices))
        self.alpha_gate = nn.Parameter(torch.zeros(self.n_choices))
        self.active_index = None
        self.inactive_index = None

    # ---- forward -------------------------------------------------------------------------------------
    def forward(self, s, pre=None, s_mask=None, pre_mask=None, rel_embed=None):
        if MixedOp.MODE in ('full', 'two'):
            order = list(self.active_index) + list(self.inactive_index)
            outs = []
            for pos, i in enumerate(order):
                if pos < len(self.active_index):
                    outs.append(self.candidate_ops[i](s, pre, s_mask, pre_mask, rel_embed))
                else:                       # inactive: forward only, cut from the graph (mixed.py:66-68)
                    with torch.no_grad():
                        outs.append(self.candidate_ops[i](s, pre, s_mask, pre_mask, rel_embed))
            gate = self.alpha_gate[torch.as_tensor(order, device=self.alpha_gate.device)]
            return MixedSumFn.apply(gate, len(self.active_index), *outs)
        if MixedOp.MODE is not None:
            raise NotImplementedError("MixedOp.MODE %r (the reference disables 'full_v2' too, mixed.py:71)" % MixedOp.MODE)
        return self.active_op(s, pre, s_mask, pre_mask, rel_embed)

    # ---- alpha bookkeeping -----------------------------------------------------------------------------
    @property
    def probs_over_ops(self):
        return F.softmax(self.alpha_prob, dim=0)

    @property
    def active_op(self):
        return self.candidate_ops[self.active_index[0]]

    @property
    def chosen_index(self):
        probs = self.probs_over_ops.data.cpu().numpy()
        index = int(np.argmax(probs))
        return index, probs[index]

    def _others(self, idx):
        return [i for i in range(self.n_choices) if i != idx]

    def set_chosen_op_active(self):
        idx, _ = self.chosen_index
        self.active_index, self.inactive_index = [idx], self._others(idx)

    def binarize(self):
        """Sample the active path from softmax(alpha_prob) and set the one-hot gate (mixed.py:131-163).
        Uses the same torch.multinomial calls as the reference so a shared seed draws the same path."""
        self.alpha_gate.data.zero_()
        probs = self.probs_over_ops
        if MixedOp.MODE == 'two':
            pair = torch.multinomial(probs.data, 2, replacement=False)
            sub = F.softmax(torch.stack([self.alpha_prob[i] for i in pair]), dim=0)
            c = torch.multinomial(sub.data, 1)[0]
            active, inactive = pair[c].item(), pair[1 - c].item()
            self.active_index, self.inactive_index = [active], [inactive]
        else:
            active = torch.multinomial(probs.data, 1)[0].item()
            self.active_index, self.inactive_index = [active], self._others(active)
        self.alpha_gate.data[active] = 1.0
        for op in self.candidate_ops:            # "avoid over-regularization": drop stale candidate grads
            if op is not None:
                for p in op.parameters():
                    p.grad = None

    def delta_ij(self, i, j):
        return 1 if i == j else 0

    def set_arch_param_grad(self):
        """alpha_prob.grad_i += sum_j gate.grad_j * p_j * (delta_ij - p_i)   (mixed.py:171-198)"""
        g = self.alpha_gate.grad.data
        if self.alpha_prob.grad is None:
            self.alpha_prob.grad = torch.zeros_like(self.alpha_prob.data)
        if MixedOp.MODE == 'two':
            involved = self.active_index + self.inactive_index
            idx = torch.as_tensor(involved, device=g.device)
            p = F.softmax(self.alpha_prob.data[idx], dim=0)
            gi = g[idx]
            self.alpha_prob.grad.data[idx] += p * (gi - (gi * p).sum())
            self.active_index = [(i, self.alpha_prob.data[i].item()) for i in self.active_index]
            self.inactive_index = [(i, self.alpha_prob.data[i].item()) for i in self.inactive_index]
        else:
            p = self.probs_over_ops.data
            self.alpha_prob.grad.data += p * (g - (g * p).sum())

    def rescale_updated_arch_param(self):
        """'two' mode only: keep logsumexp of the two touched alphas unchanged (mixed.py:200-208)."""
        pairs = self.active_index + self.inactive_index
        idx = [i for i, _ in pairs]
        old = [a for _, a in pairs]
        new = [self.alpha_prob.data[i].item() for i in idx]
        offset = math.log(sum(math.exp(a) for a in new) / sum(math.exp(a) for a in old))
        for i in idx:
            self.alpha_prob.data[i] -= offset
