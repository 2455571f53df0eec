"""ITM hard-negative mining forward (train_itm.py:299-363) at the reference's size: NEG_BATCHSIZE x NEG_RANDSIZE = 50 x 64 =
3 200 (image, caption) pairs per forward, one side repeated 64 times.  score_pairs() (every unique image / caption encoded
once) against the expanded forward the reference runs; inference mode, bf16 arm."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmnas_b200
from mmnas_b200 import genotypes
from mmnas_b200.data.synthetic import Cfg, make_batch, init_dict, spec_for
from mmnas_b200.model.nets import Net_Full

DEV = 'cuda'
torch.manual_seed(888)
anchors, group = 50, 64
P = anchors * group
spec = spec_for('itm', batch=(P + 2) // 3)
cfg = Cfg(genotype=genotypes.shipped('mmnas_itm'))
(frcn, bbox, rel, caps, rel_cap), _ = make_batch(spec, seed=5)
frcn, bbox, rel, caps, rel_cap = (t[:P].to(DEV) for t in (frcn, bbox, rel, caps, rel_cap))
net = Net_Full(cfg, init_dict(spec), task='itm').to(DEV).eval()
anchor = torch.arange(anchors, device=DEV).repeat_interleave(group)
other = torch.randperm(P, device=DEV)
res = {}
flops_pair = 241e9 / 64        # 241 GF backbone forward per 64 ITM pairs
for repeated in ('image', 'caption'):
    if repeated == 'image':
        img_index, cap_index = anchor, other
        images, captions = (frcn[:anchors], bbox[:anchors], rel[:anchors]), (caps, rel_cap)
    else:
        img_index, cap_index = other, anchor
        images, captions = (frcn, bbox, rel), (caps[:anchors], rel_cap[:anchors])
    expanded = (images[0][img_index], images[1][img_index], images[2][img_index], captions[0][cap_index], captions[1][cap_index])
    row = {}
    with mmnas_b200.precision('bf16'), torch.no_grad():
        for name, fn in (('expanded_forward', lambda: net(expanded)), ('score_pairs', lambda: net.score_pairs(images, captions, img_index, cap_index))):
            for _ in range(2): out = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): out = fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            row[name] = {'ms': ms, 'pairs_per_s': P / ms * 1e3, 'backbone_tflops_expanded_equivalent': flops_pair * P / ms / 1e9}
            row[name + '_scores'] = out
    a, b = row.pop('expanded_forward_scores'), row.pop('score_pairs_scores')
    row['max_abs_score_diff'] = float((a - b).abs().max())
    res['repeated_' + repeated] = row
    print(repeated, {k: (round(v['ms'], 2), int(v['pairs_per_s'])) if isinstance(v, dict) else v for k, v in row.items()})
if len(sys.argv) > 1:
    json.dump({'pairs_per_forward': P, 'anchors': anchors, 'group': group, **res}, open(sys.argv[1], 'w'), indent=1)
