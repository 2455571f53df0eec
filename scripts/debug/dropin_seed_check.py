"""How the batch seeds of tests/_dropin_worker.py sections 3-5 were chosen: the UNMODIFIED reference (baseline/_ref) on the
CPU in float32 against itself in float64.  A seed is usable when the two agree to ~1e-5 on every gradient tensor, i.e. no
ReLU pre-activation lies within float32 rounding of zero (such a unit takes either branch depending on the summation
order of whoever evaluates it, and moves one token's contribution to a weight gradient by 1e-3 .. 1e-2).

    python scripts/debug/dropin_seed_check.py [vgd_seed itm_seed]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, 'baseline', '_ref'))
from mmnas_b200.data.synthetic import Cfg, make_batch, init_dict          # noqa: E402
from tests.util import condition_rsa_, normwise                           # noqa: E402
from tests import _dropin_worker as W                                     # noqa: E402


def main():
    seeds = tuple(int(a) for a in sys.argv[1:3]) if len(sys.argv) >= 3 else (888, 892)
    runs = {}
    for dt in (torch.float32, torch.float64):
        res = {}
        W.search_weight_steps(res, Cfg, make_batch, init_dict, condition_rsa_, 'cpu', dtype=dt, seeds=seeds)
        runs[dt] = res
    for key in runs[torch.float32]:
        a, b = runs[torch.float32][key], runs[torch.float64][key]
        floor = 1e-2 * max(float(g.abs().max()) for g in b['grads'].values())
        rows = sorted(((normwise(a['grads'][n], g, floor), n) for n, g in b['grads'].items()), reverse=True)
        print(key, 'seed', seeds[0 if key.endswith('vgd') else 1], 'pred', '%.2e' % normwise(a['pred'], b['pred']),
              'worst gradients:', [('%.2e' % e, n) for e, n in rows[:3]])


if __name__ == '__main__':
    main()
