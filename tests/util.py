"""Shared helpers for the test-suite (golden loading, the normwise metric of SURVEY §8c)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name, group=None):
    z = np.load(os.path.join(GOLDEN, name))
    out = {}
    for k in z.files:
        if group is not None:
            if not k.startswith(group + '/'):
                continue
            kk = k[len(group) + 1:]
        else:
            kk = k
        a = z[k]
        out[kk] = a if a.dtype.kind in 'US' else torch.from_numpy(np.array(a))
    return out


def params_of(rec, prefix='p.'):
    return {k[len(prefix):]: v for k, v in rec.items() if k.startswith(prefix)}


def literal(rec, key):
    return ast.literal_eval(str(rec[key]))


def normwise(new, ref, floor=0.0):
    """max|new-ref| / max(max|ref|, floor)  — the tolerance metric calibrated in SURVEY.md §8(c).

    ``floor`` guards tensors whose true value is zero (e.g. the gradient of a bias that feeds a
    softmax over the same axis): pass a small fraction of the largest gradient in the model."""
    new = new.detach().double().cpu()
    ref = ref.detach().double().cpu()
    den = max(ref.abs().max().item(), floor)
    num = (new - ref).abs().max().item()
    if den == 0:
        return num
    return num / den


def grad_floor(rec, frac=1e-2, prefix='g.'):
    """frac x the largest reference gradient entry over all parameters of a golden record."""
    return frac * max(v.abs().max().item() for k, v in rec.items() if k.startswith(prefix))


class Parity:
    """Collects normwise errors of many tensors, logs them (gpurun_out/parity.jsonl when that directory exists)
    and fails once at the end listing every tensor over tolerance."""

    def __init__(self, label):
        self.label, self.rows = label, []

    def add(self, name, new, ref, tol, floor=0.0):
        self.rows.append((name, normwise(new, ref, floor), tol))

    def check(self):
        import json
        worst = sorted(self.rows, key=lambda r: -r[1] / r[2])[:5]
        out = os.path.join(os.path.dirname(GOLDEN), '..', 'gpurun_out')
        if os.path.isdir(out):
            with open(os.path.join(out, 'parity.jsonl'), 'a') as f:
                f.write(json.dumps({'case': self.label, 'n': len(self.rows),
                                    'worst': [(n, float('%.3g' % e), t) for n, e, t in worst]}) + '\n')
        bad = [(n, float('%.3g' % e), t) for n, e, t in self.rows if not e < t]
        assert not bad, '%s: over tolerance: %s' % (self.label, bad)
