"""Callers of the operator hot path: cells, backbones and the task nets (train-time `Net_Full`, search-time
`Net_Search`), mirroring mmnas/model/full_{vqa,vgd,itm}.py and hygr_{vqa,vgd,itm}.py.

Only the 30-node backbone is the CUDA hot path; stem (embedding, LSTM, imgfeat_linear, masks) and task heads
(AttFlat, projections) are the adjacent rows of SURVEY §8f and stay on PyTorch for now.  Module names,
creation order (=> identical default init under one seed) and state-dict keys follow the reference, incl. its
`backnone` spelling, so reference checkpoints load with load_state_dict.

The one deliberate difference: RSA blocks receive a `RelGeometry` handle (raw 4-d geometry + the
`linear_y_rel` layer) instead of the dense relu(linear_y_rel(g)) [B,N,N,64] tensor, unless
`rel_mode='dense'` is requested.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import runtime
from ..functional import StemImageFn, lstm
from .mixed import MixedOp
from .modules import AttFlat, LayerNorm, RelGeometry, linear
from ..functional import AddLayerNormFn
from ..utils.ops_adapter import OpsAdapter

OPS_ADAPTER = OpsAdapter()


def _chain(rows, s, pre, s_mask, pre_mask, rel_embed):
    """Cell forward (full_vqa.py:24-28 / hygr_vqa.py:23-27): every node is the sum over its row of ops; the
    single-op rows of every shipped arch skip the `0 + x` add so the bf16 shadow of x survives."""
    for ops in rows:
        if len(ops) == 1:
            s = ops[0](s, pre, s_mask, pre_mask, rel_embed)
        else:
            s = sum(op(s, pre, s_mask, pre_mask, rel_embed) for op in ops)
    return s


class Cell_Full(nn.Module):
    def __init__(self, __C, type):
        super().__init__()
        self.dag = nn.ModuleList(
            nn.ModuleList(OPS_ADAPTER.OPS[name](__C, norm=__C.OPS_NORM, residual=__C.OPS_RESIDUAL) for name in node)
            for node in __C.GENOTYPE[type])
        self.NODES = len(self.dag)

    def forward(self, s, pre=None, s_mask=None, pre_mask=None, rel_embed=None):
        return _chain(self.dag, s, pre, s_mask, pre_mask, rel_embed)


class Cell_Search(nn.Module):
    def __init__(self, __C, type):
        super().__init__()
        self.dag = nn.ModuleList(nn.ModuleList([MixedOp(__C, type + '_safe')]) for _ in range(__C.NODES[type]))

    def forward(self, s, pre=None, s_mask=None, pre_mask=None, rel_embed=None):
        return _chain(self.dag, s, pre, s_mask, pre_mask, rel_embed)


class _Backbone(nn.Module):
    CELL = None

    def __init__(self, __C):
        super().__init__()
        self.cells_enc = nn.ModuleList(self.CELL(__C, type='enc') for _ in range(__C.LAYERS))
        self.cells_dec = nn.ModuleList(self.CELL(__C, type='dec') for _ in range(__C.LAYERS))

    def forward(self, x, y, x_mask, y_mask, x_rel_embed, y_rel_embed):
        for cell in self.cells_enc:
            x = cell(s=x, s_mask=x_mask, rel_embed=x_rel_embed)
        for cell in self.cells_dec:
            y = cell(s=y, pre=x, s_mask=y_mask, pre_mask=x_mask, rel_embed=y_rel_embed)
        return x, y


class Backbone_Full(_Backbone):
    CELL = Cell_Full


class Backbone_Search(_Backbone):
    CELL = Cell_Search


def make_mask(feature):
    return (torch.sum(torch.abs(feature), dim=-1) == 0).unsqueeze(1).unsqueeze(2)


class _NetBase(nn.Module):
    """Stem shared by the three tasks and by Net_Full / Net_Search."""
    BACKBONE = None
    SEARCH = False

    def __init__(self, __C, init_dict, task='vqa', rel_mode='geometry'):
        super().__init__()
        assert task in ('vqa', 'vgd', 'itm') and rel_mode in ('geometry', 'dense')
        self.task, self.rel_mode = task, rel_mode
        self.BBOX_FEATURE = __C.BBOX_FEATURE
        self.SCORES_LOSS = getattr(__C, 'SCORES_LOSS', None)
        self.embedding = nn.Embedding(num_embeddings=init_dict['token_size'], embedding_dim=__C.WORD_EMBED_SIZE)
        self.embedding.weight.data.copy_(torch.from_numpy(init_dict['pretrained_emb']))
        self.lstm = nn.LSTM(input_size=__C.WORD_EMBED_SIZE, hidden_size=__C.HSIZE, num_layers=1, batch_first=True)
        feat = __C.FRCNFEAT_SIZE
        if __C.BBOX_FEATURE:
            self.bboxfeat_linear = nn.Linear(5, __C.BBOXFEAT_EMB_SIZE)
            feat += __C.BBOXFEAT_EMB_SIZE
        self.imgfeat_linear = nn.Linear(feat, __C.HSIZE)
        if task == 'itm' and not self.SEARCH:       # full_itm.py:75 creates it before the backbone (same-seed init parity)
            self.linear_y_rel = nn.Linear(4, __C.REL_SIZE)
        self.backnone = self.BACKBONE(__C)
        self.attflat_x = AttFlat(__C)
        if task == 'vgd':       # full_vgd.py:83-87
            self.attfc_y = nn.Linear(__C.HSIZE, __C.ATTFLAT_OUT_SIZE)
            self.proj_norm = LayerNorm(__C.ATTFLAT_OUT_SIZE)
            self.proj_scores = nn.Linear(__C.ATTFLAT_OUT_SIZE, 1)
            self.proj_reg = nn.Linear(__C.ATTFLAT_OUT_SIZE, 4)
        else:                   # full_vqa.py:78-80, full_itm.py (proj -> 1 logit)
            self.attflat_y = AttFlat(__C)
            self.proj_norm = LayerNorm(__C.ATTFLAT_OUT_SIZE)
            self.proj = nn.Linear(__C.ATTFLAT_OUT_SIZE, init_dict['ans_size'] if task == 'vqa' else 1)
        if self.SEARCH:
            self.linear_x_rel = nn.Linear(3, __C.REL_SIZE)
        if not hasattr(self, 'linear_y_rel'):
            self.linear_y_rel = nn.Linear(4, __C.REL_SIZE)

    def stem(self, input):
        """Everything in front of the backbone (full_vqa.py:90-103): question embedding + LSTM, region projection, the
        two padding masks, and the pairwise box geometry when the loader ships raw boxes."""
        frcn_feat, bbox_feat, y_rel, ques_ix, x_rel = input
        x_mask = make_mask(ques_ix.unsqueeze(2))
        x_in = lstm(self.embedding(ques_ix), self.lstm)
        if self.BBOX_FEATURE:       # not used by any shipped config (BBOX_FEATURE = False, train_vqa.py:136): torch path
            y_mask = make_mask(frcn_feat)
            frcn_feat = torch.cat((frcn_feat, self.bboxfeat_linear(bbox_feat)), dim=-1)
            y_in = self.imgfeat_linear(frcn_feat)
        else:                       # mask + bf16 cast in one pass over the features, projection on our GEMM
            y_in, y_mask = StemImageFn.apply(frcn_feat, self.imgfeat_linear.weight, self.imgfeat_linear.bias,
                                             runtime.get_precision())
        # x_rel is never consumed (no relation op is an encoder candidate); the reference still embeds it in
        # Net_Search (hygr_vqa.py:130) — the parameters exist here for checkpoint parity, the dead matmul does not run.
        if y_rel.dim() == 3:
            # compact loader format: raw boxes [B,N,4] instead of the [B,N,N,4] log-geometry the
            # reference's DataLoader computes per sample on the CPU (load_data_vqa.py:7-33,236-239) — built here
            from .. import kernels as K
            from .._lib import require_cuda
            require_cuda(y_rel)
            pad = y_mask.reshape(y_mask.shape[0], -1).contiguous().view(torch.uint8)
            y_rel = K.box_geometry(y_rel.contiguous().float(), pad)
        return x_in, y_in, x_mask, y_mask, x_rel, y_rel

    def forward(self, input):
        x_in, y_in, x_mask, y_mask, x_rel, y_rel = self.stem(input)
        ex = getattr(self, '_executor', None)
        if ex is not None and ex.usable(self.rel_mode):
            # engine search step: the whole supernet backbone is one autograd node over static per-candidate plans
            from ..executor import BackboneFn
            x_out, y_out = BackboneFn.apply(x_in, y_in, x_mask, y_mask, y_rel, ex)
            return self.head(x_out, y_out, x_mask, y_mask)
        if self.rel_mode == 'geometry':
            y_rel = RelGeometry(y_rel, self.linear_y_rel)
        else:
            y_rel = F.relu(self.linear_y_rel(y_rel))
        x_out, y_out = self.backnone(x_in, y_in, x_mask, y_mask, x_rel, y_rel)
        return self.head(x_out, y_out, x_mask, y_mask)

    def _add_norm(self, a, b):
        """proj_norm(a + b): one residual + LayerNorm kernel on CUDA."""
        if b.is_cuda:
            return AddLayerNormFn.apply(a, b, self.proj_norm.a_2, self.proj_norm.b_2, self.proj_norm.eps)
        return self.proj_norm(a + b)

    def head(self, x_out, y_out, x_mask, y_mask):
        if self.task == 'vgd':      # full_vgd.py:105-112
            xy = self._add_norm(self.attflat_x(x_out, x_mask).unsqueeze(1), linear(y_out, self.attfc_y))
            scores = linear(xy, self.proj_scores).squeeze(-1)
            if self.SCORES_LOSS == 'kld':
                scores = F.log_softmax(scores, dim=-1)
            return scores, linear(xy, self.proj_reg)
        xy = self._add_norm(self.attflat_x(x_out, x_mask), self.attflat_y(y_out, y_mask))
        out = linear(xy, self.proj)
        if self.task == 'itm':      # full_itm.py:109-110
            out = torch.sigmoid(out.squeeze(-1))
        return out

    make_mask = staticmethod(make_mask)

    # ---- ITM hard-negative mining / retrieval scoring ----------------------------------------------
    @torch.no_grad()
    def score_pairs(self, images, captions, img_index, cap_index):
        """Matching scores of the pairs (images[img_index[p]], captions[cap_index[p]]) in inference mode, what
        train_itm.py:307-320 / :340-353 (hard-negative mining: NEG_BATCHSIZE x NEG_RANDSIZE = 50 x 64 = 3 200 pairs per
        forward) and :476-500 (retrieval evaluation) compute with `net(input)` on inputs where one side is REPEATED
        64 times.  Here every unique image and caption is encoded once:
          * captions: embedding, LSTM and the whole encoder (it never sees the image) run on the unique captions;
          * images: the stem projection, the padding mask, the pairwise geometry and the decoder blocks that precede
            the first GuidedAtt block run on the unique images;
        then both sides are gathered to the pair list and the rest of the decoder + the head run on the pairs.  Same
        result as `self(expanded inputs)` (tests/test_gpu_nets.py), with the repeated side's work divided by the
        repetition count.  images = (frcn_feat [U,N,F], bbox_feat, rel_img [U,N,N,4] or boxes [U,N,4]); captions =
        (cap_ix [V,T], rel_cap); img_index / cap_index: int64 [P] on the same device."""
        frcn_feat, bbox_feat, y_rel = images
        cap_ix, x_rel = captions
        x_mask_u = make_mask(cap_ix.unsqueeze(2))
        x_u = lstm(self.embedding(cap_ix), self.lstm)
        y_u, y_mask_u = StemImageFn.apply(frcn_feat, self.imgfeat_linear.weight, self.imgfeat_linear.bias,
                                          runtime.get_precision())
        if y_rel.dim() == 3:
            from .. import kernels as K
            y_rel = K.box_geometry(y_rel.contiguous().float(), y_mask_u.reshape(y_mask_u.shape[0], -1).contiguous().view(torch.uint8))
        geo_u = RelGeometry(y_rel, self.linear_y_rel) if self.rel_mode == 'geometry' else F.relu(self.linear_y_rel(y_rel))
        for cell in self.backnone.cells_enc:
            x_u = cell(s=x_u, s_mask=x_mask_u, rel_embed=x_rel)
        rows = [ops for cell in self.backnone.cells_dec for ops in cell.dag]
        from .modules import GuidedAtt
        n_free = 0                               # leading decoder nodes that do not read the encoder output
        while n_free < len(rows) and not any(isinstance(op, (GuidedAtt, MixedOp)) for op in rows[n_free]):
            n_free += 1
        y_u = _chain(rows[:n_free], y_u, None, y_mask_u, None, geo_u)
        x, x_mask = x_u[cap_index], x_mask_u[cap_index]
        y, y_mask = y_u[img_index], y_mask_u[img_index]
        if isinstance(geo_u, RelGeometry):
            geo = RelGeometry(geo_u.g4[img_index], self.linear_y_rel)
        else:
            geo = geo_u[img_index]
        y = _chain(rows[n_free:], y, x, y_mask, x_mask, geo)
        return self.head(x, y, x_mask, y_mask)

    @staticmethod
    def hard_negatives(scores, neg_idx_list, group, k):
        """train_itm.py:316-320: per anchor, the k highest-scoring of its `group` random negatives."""
        scores = scores.view(-1, group)
        top = torch.argsort(scores, dim=-1, descending=True)[:, :k]
        rows = torch.arange(top.size(0), device=top.device).unsqueeze(1).expand_as(top)
        return neg_idx_list.to(top.device)[rows, top]


class Net_Full(_NetBase):
    """Train-time net built from a genotype (arch/*.json entry in __C.GENOTYPE)."""
    BACKBONE = Backbone_Full


class Net_Search(_NetBase):
    """Search-time supernet: one MixedOp per node + the architecture-parameter bookkeeping of hygr_vqa.py:124-297."""
    BACKBONE = Backbone_Search
    SEARCH = True
    # MCAN-like prior written over the alpha init (hygr_vqa.py:142-156)
    PRIOR_ENC = ['self_att_64', 'feed_forward'] * 6
    PRIOR_DEC = ['rel_self_att_64', 'guided_att_64', 'feed_forward'] * 7

    def __init__(self, __C, init_dict, task='vqa', rel_mode='geometry'):
        super().__init__(__C, init_dict, task, rel_mode)
        self._alpha_init_type = __C.ALPHA_INIT_TYPE
        self._redundant_modules = None
        self._unused_modules = None
        self.init_arch()
        self._net_weights = [(n, p) for n, p in self.named_parameters()
                             if 'alpha_prob' not in n and 'alpha_gate' not in n]

    # ---- architecture parameters ---------------------------------------------------------------------
    def init_arch(self):
        self._alphas_prob = [(n, p) for n, p in self.named_parameters() if 'alpha_prob' in n]
        self._alphas_gate = [(n, p) for n, p in self.named_parameters() if 'alpha_gate' in n]
        for _, p in self._alphas_prob:
            if self._alpha_init_type == 'normal':
                p.data.normal_(0, 1e-3)
            elif self._alpha_init_type == 'uniform':
                p.data.uniform_(-1e-3, 1e-3)
        # The reference then overwrites the first 12 alphas with an encoder-sized prior and the rest with a
        # decoder-sized one (so it assumes 12 encoder nodes, hygr_vqa.py:148); the same rule is applied here.
        prior = self.PRIOR_ENC + self.PRIOR_DEC
        for ix, (name, (_, p)) in enumerate(zip(prior, self._alphas_prob)):
            cands = OPS_ADAPTER.Used_OPS['enc_safe' if ix < 12 else 'dec_safe']
            init = np.full(len(cands), -1., dtype=np.float32)
            init[cands.index(name)] = 1.
            p.data = torch.from_numpy(init).to(p.device)

    @property
    def redundant_modules(self):
        if self._redundant_modules is None:
            self._redundant_modules = [m for m in self.modules() if str(m).startswith('MixedOp')]
        return self._redundant_modules

    def reset_binary_gates(self, batched=False, group=None):
        """Sample every node's active path (MixedOp.binarize, mixed.py:131-163).

        batched=True (the step harness) keeps the reference's random-number consumption.  The reference draws
        `torch.multinomial(probs, 1)` once per MixedOp, in module registration order (hygr_vqa.py:168-172 over
        mixed.py:151).  For a single draw torch implements that as  q = empty_like(probs).exponential_(1);
        argmax(probs / q)  (ATen multinomial_out, the "gumbel" fast path, CPU and CUDA alike) — the generator is
        consumed by exactly one exponential_ call on a [K] tensor.  Here the SAME exponential_ calls are issued in the
        same order (so one seed draws the same path as the reference: tests/test_host_logic.py checks five consecutive
        draws against the unmodified reference), but everything around them is batched: one softmax and one argmax per
        node width instead of seven tiny kernels per node, ONE host read of all 30 picks instead of 30 `.item()`
        syncs, one multi-tensor copy for the one-hot gates.  The candidates' .grad buffers are left alone (the harness
        zeroes the flat gradient buffer, which is what the reference's dummy-loss terms turn the None grads into).
        Under data parallelism the picks are broadcast from rank 0, so every rank runs the same path by construction
        (the reference relies on identically seeded ranks).  Not for MODE 'two'."""
        if not batched:
            self.__dict__.pop('_presampled', None)
            for m in self.redundant_modules:
                m.binarize()
            return
        pre = self.__dict__.pop('_presampled', None)
        if pre is not None and self._presample_valid(pre, group):
            pre['done'].synchronize()                                  # normally long complete: drawn a step ago
            torch.cuda.current_stream(pre['picks'].device).wait_event(pre['done'])
            pre['picks'].record_stream(torch.cuda.current_stream(pre['picks'].device))
            self._apply_picks(pre['picks'], pre['host'].tolist())
            return
        picks = self._draw_picks(group)
        self._apply_picks(picks, picks.tolist())

    def _sampling_plan(self):
        mods = self.redundant_modules
        plan = getattr(self, '_sample_plan', None)
        if plan is None or plan[0] != len(mods):
            groups = {}
            for i, m in enumerate(mods):
                groups.setdefault(m.n_choices, []).append(i)
            order = [i for k in groups for i in groups[k]]             # module index of every row of the stacked picks
            plan = self._sample_plan = (len(mods), groups, order)
        return plan

    def _draw_picks(self, group):
        """Device side of the batched draw on the current stream: the per-node exponential_ calls in registration
        order, one softmax + argmax per node width, the broadcast from rank 0.  Returns the picks in plan order."""
        mods = self.redundant_modules
        _, groups, order = self._sampling_plan()
        with torch.no_grad():
            dev = mods[0].alpha_prob.device
            qs = [torch.empty(m.n_choices, dtype=m.alpha_prob.dtype, device=dev).exponential_(1) for m in mods]
            picks = []
            for k, idxs in groups.items():
                probs = F.softmax(torch.stack([mods[i].alpha_prob.data for i in idxs]), dim=1)
                picks.append(torch.argmax(probs / torch.stack([qs[i] for i in idxs]), dim=1))
            picks = torch.cat(picks)                                   # in `order`
            if group is not False and torch.distributed.is_available() and torch.distributed.is_initialized() \
                    and torch.distributed.get_world_size(group) > 1:
                torch.distributed.broadcast(picks, 0, group=group)
        return picks

    def _apply_picks(self, picks, picks_list):
        """One-hot gates on the device (current stream) and the host-side active / inactive bookkeeping."""
        mods = self.redundant_modules
        _, groups, order = self._sampling_plan()
        with torch.no_grad():
            off, onehots, gates = 0, [], []
            for k, idxs in groups.items():
                oh = F.one_hot(picks[off:off + len(idxs)], k).to(mods[idxs[0]].alpha_gate.dtype)
                off += len(idxs)
                gates += [mods[i].alpha_gate.data for i in idxs]
                onehots += list(oh.unbind(0))
            torch._foreach_copy_(gates, onehots)
        for i, a in zip(order, picks_list):
            m = mods[i]
            object.__setattr__(m, 'active_index', [a])                 # plain attributes: skip nn.Module.__setattr__
            object.__setattr__(m, 'inactive_index', [j for j in range(m.n_choices) if j != a])

    # ---- drawing one step ahead -------------------------------------------------------------------------------------
    # The picks have to reach the host before the step can be launched (they choose which kernels run), and a device ->
    # host read at the top of a step drains the launch queue: the GPU idles while the host catches up.  The architecture
    # parameters only change in the architecture step, so the NEXT step's draw can be issued as soon as this step is
    # launched, on its own stream, ordered only after the last alpha update; by the time the next step starts the
    # picks are already in pinned host memory.  Same generator, same calls, same order as drawing at the top of the step.
    def alphas_updated(self):
        """Called by the step harness right after the architecture optimizer's update was launched."""
        dev = self.redundant_modules[0].alpha_prob.device
        if dev.type == 'cuda':
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self.__dict__['_alpha_event'] = ev

    def _generator_state(self, dev):
        return torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()].get_state()

    def _alpha_versions(self):
        return tuple((m.alpha_prob._version, m.alpha_prob.data_ptr()) for m in self.redundant_modules)

    def presample(self, group=None):
        dev = self.redundant_modules[0].alpha_prob.device
        if dev.type != 'cuda':
            return
        st = self.__dict__.get('_sample_stream')
        if st is None:
            st = self.__dict__['_sample_stream'] = torch.cuda.Stream(device=dev)
            n = len(self.redundant_modules)
            self.__dict__['_picks_host'] = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(2)]
            self.__dict__['_picks_flip'] = 0
        ev = self.__dict__.get('_alpha_event')
        if ev is not None:
            st.wait_event(ev)
        self.__dict__['_picks_flip'] ^= 1
        host = self.__dict__['_picks_host'][self.__dict__['_picks_flip']]
        with torch.cuda.stream(st):
            picks = self._draw_picks(group)
            host.copy_(picks, non_blocking=True)
            done = torch.cuda.Event()
            done.record(st)
        self.__dict__['_presampled'] = {'picks': picks, 'host': host, 'done': done, 'group': group,
                                        'gen': self._generator_state(dev), 'alphas': self._alpha_versions()}

    def _presample_valid(self, pre, group):
        dev = pre['picks'].device
        return (pre['group'] is group and pre['alphas'] == self._alpha_versions()
                and torch.equal(pre['gen'], self._generator_state(dev)))

    def unused_modules_off(self):
        self._unused_modules = []
        for m in self.redundant_modules:
            involved = m.active_index + m.inactive_index if MixedOp.MODE in ('full', 'two', 'full_v2') else m.active_index
            unused = {}
            for i in range(m.n_choices):
                if i not in involved:
                    unused[i] = m.candidate_ops[i]
                    m.candidate_ops[i] = None
            self._unused_modules.append(unused)

    def unused_modules_back(self):
        if self._unused_modules is None:
            return
        for m, unused in zip(self.redundant_modules, self._unused_modules):
            for i, op in unused.items():
                m.candidate_ops[i] = op
        self._unused_modules = None

    def set_arch_param_grad(self, batched=False):
        """alpha_prob.grad_i += sum_j gate.grad_j p_j (delta_ij - p_i) for every node (MixedOp.set_arch_param_grad,
        mixed.py:171-198).  batched=True ('full' mode only) evaluates the rule for all nodes of one width at once —
        a dozen small kernels instead of ~180."""
        if not batched or MixedOp.MODE == 'two':
            for m in self.redundant_modules:
                m.set_arch_param_grad()
            return
        groups = {}
        for m in self.redundant_modules:
            groups.setdefault(m.n_choices, []).append(m)
        with torch.no_grad():
            for k, ms in groups.items():
                for m in ms:
                    if m.alpha_prob.grad is None:
                        m.alpha_prob.grad = torch.zeros_like(m.alpha_prob.data)
                p = F.softmax(torch.stack([m.alpha_prob.data for m in ms]), dim=1)
                g = torch.stack([m.alpha_gate.grad for m in ms])
                upd = p * (g - (g * p).sum(1, keepdim=True))
                torch._foreach_add_([m.alpha_prob.grad for m in ms], list(upd.unbind(0)))

    def rescale_updated_arch_param(self):
        for m in self.redundant_modules:
            m.rescale_updated_arch_param()

    def set_chosen_op_active(self):
        for m in self.redundant_modules:
            m.set_chosen_op_active()

    def alpha_prob_parameters(self):
        return (p for _, p in self._alphas_prob)

    def alpha_gate_parameters(self):
        return (p for _, p in self._alphas_gate)

    def named_alpha_prob_parameters(self):
        return iter(self._alphas_prob)

    def named_alpha_gate_parameters(self):
        return iter(self._alphas_gate)

    def net_parameters(self):
        return (p for _, p in self._net_weights)

    def named_net_parameters(self):
        return iter(self._net_weights)

    # ---- genotype export (format of arch/*.json entries) -----------------------------------------------
    def _alphas_of(self, kind):
        return [p for n, p in self._alphas_prob if kind in n]

    def parse(self, alpha_param_list, type):
        return [[OPS_ADAPTER.Used_OPS[type][int(torch.topk(a, 1)[1][0])]] for a in alpha_param_list]

    def genotype(self):
        return {'enc': self.parse(self._alphas_of('enc'), 'enc'), 'dec': self.parse(self._alphas_of('dec'), 'dec')}

    def parse_weights(self, alpha):
        with torch.no_grad():
            return [F.softmax(a, dim=-1).data.cpu().numpy() for a in alpha]

    def genotype_weights(self):
        return {'w_enc': self.parse_weights(self._alphas_of('enc')), 'w_dec': self.parse_weights(self._alphas_of('dec'))}
