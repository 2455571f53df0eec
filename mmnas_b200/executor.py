"""Static-plan executor of the supernet backbone for the search step (search_vqa.py:278-337).

The sampled path changes every step, so the search step cannot be replayed from ONE CUDA graph, and in eager mode it
was bound by the host: 30 autograd nodes per direction, each allocating its outputs and workspace and marshalling a
descriptor.  Here everything that does not depend on the sample is prepared once:

  * every node of the 12 + 18 node chain owns static activation / gradient buffers (node i reads buffer i and writes
    buffer i + 1; the backward reads gradient buffer i + 1 and writes gradient buffer i);
  * every CANDIDATE of every node owns a fully filled block descriptor (mmnas_att_block / mmnas_ffn_block: weights,
    bf16 weight copies, gradient destinations inside the flat gradient buffer, dropout site, workspace) — running a
    candidate is one foreign call with no per-call setup;
  * the whole backbone is ONE autograd node (BackboneFn): forward = 30 block calls (weight step, MixedOp.MODE None:
    the sampled candidate of each node) or all 96 candidates + one fused weighted sum per node (architecture step,
    MODE 'full', mixed.py:60-68); backward = the reverse chain, the six-to-eleven GuidedAtt blocks adding their
    key/value gradients into ONE encoder-output gradient buffer through the dgrad GEMM's accumulate epilogue, and in
    'full' mode one pass per node that yields every alpha_gate gradient <o_k, dOut> and the active candidate's
    gradient (mmnas_mixed_alpha_dot).

Semantics are those of Cell_Search / MixedOp (hygr_vqa.py:23-27, mixed.py:59-106) through the same kernels as the
module path; tests/test_gpu_nets.py checks the executor against the per-module autograd path and the float64 oracle.
Only engine steps use it (gradients accumulate straight into engine.FlatGrads; runtime.direct_grads must be on)."""
import ctypes

import torch
from torch.autograd import Function

from . import _lib, runtime
from . import kernels as K
from .model import modules as M
from .model.mixed import MixedOp


class _Cand:
    """One candidate block of one node: descriptor + entry points + the parameters whose gradients it produces."""
    __slots__ = ('kind', 'desc', 'ref', 'fwd', 'bwd', 'params', 'ws', 'guided', 'rel')


def _grad_ptr(p):
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous():
        raise RuntimeError('executor: parameter without a flat fp32 gradient view (build engine.FlatGrads first)')
    return g.data_ptr()


class SearchExecutor:
    def __init__(self, net):
        self.net = net
        self.key = None
        self.nodes_x = [row[0] for cell in net.backnone.cells_enc for row in cell.dag]
        self.nodes_y = [row[0] for cell in net.backnone.cells_dec for row in cell.dag]
        for m in self.nodes_x + self.nodes_y:
            if not isinstance(m, MixedOp):
                raise TypeError('SearchExecutor needs a Net_Search backbone (one MixedOp per node)')

    # ------------------------------------------------------------------------------------------------ planning
    def usable(self, rel_mode):
        return runtime.direct_grads and rel_mode == 'geometry' and MixedOp.MODE in (None, 'full') and self.net.training is not None

    def _build(self, B, Nx, Ny, H, dev, bf):
        lib = _lib.load()
        f32 = dict(dtype=torch.float32, device=dev)
        nx, ny = len(self.nodes_x), len(self.nodes_y)
        self.bf = bf
        self.xs = [torch.empty((B, Nx, H), **f32) for _ in range(nx + 1)]
        self.ys = [torch.empty((B, Ny, H), **f32) for _ in range(ny + 1)]
        self.dxs = [torch.empty((B, Nx, H), **f32) for _ in range(nx + 1)]
        self.dys = [torch.empty((B, Ny, H), **f32) for _ in range(ny + 1)]
        b16 = dict(dtype=torch.bfloat16, device=dev)
        self.x16s = [torch.empty((B, Nx, H), **b16) if bf else None for _ in range(nx + 1)]
        self.y16s = [torch.empty((B, Ny, H), **b16) if bf else None for _ in range(ny + 1)]
        self.xmask = torch.zeros((B, Nx), dtype=torch.uint8, device=dev)
        self.ymask = torch.zeros((B, Ny), dtype=torch.uint8, device=dev)
        self.g4 = torch.empty((B, Ny, Ny, 4), **f32)
        self.dcand_x = torch.empty((B, Nx, H), **f32)
        self.dcand_y = torch.empty((B, Ny, H), **f32)
        self.cand_out = {}                         # (side, node, cand) -> fp32 output buffer, 'full' mode only (lazy)
        rng = runtime.rng_state(dev)
        self.rng = rng
        lin = self.net.linear_y_rel
        self.geo_params = (lin.weight, lin.bias)
        max_bwd = 256
        self.cands_x, self.cands_y = [], []
        for side, nodes, cands, N in (('x', self.nodes_x, self.cands_x, Nx), ('y', self.nodes_y, self.cands_y, Ny)):
            for i, node in enumerate(nodes):
                row = []
                for c, op in enumerate(node.candidate_ops):
                    cd = self._plan_candidate(op, side, i, B, N, Nx, H, dev, bf, lib)
                    fwd, bwd = _lib.workspace_bytes(cd.desc)
                    cd.ws = torch.empty(fwd, dtype=torch.uint8, device=dev)
                    cd.desc.workspace = cd.ws.data_ptr()
                    max_bwd = max(max_bwd, bwd)
                    row.append(cd)
                cands.append(row)
        self.bws = torch.empty(max_bwd, dtype=torch.uint8, device=dev)     # backward scratch: blocks run one at a time
        for row in self.cands_x + self.cands_y:
            for cd in row:
                cd.desc.bwd_workspace = self.bws.data_ptr()
        self.gates_x = [m.alpha_gate for m in self.nodes_x]
        self.gates_y = [m.alpha_gate for m in self.nodes_y]

    def _plan_candidate(self, op, side, i, B, N, Nx, H, dev, bf, lib):
        cd = _Cand()
        xs, x16s, dxs = (self.xs, self.x16s, self.dxs) if side == 'x' else (self.ys, self.y16s, self.dys)
        training = self.net.training
        if isinstance(op, M.FeedForward):
            d = _lib.FfnBlock()
            cd.kind, cd.guided, cd.rel = 'ffn', False, False
            w1, w2 = op.mlp.fc.linear, op.mlp.linear
            d.precision = 1 if bf else 0
            d.M, d.H, d.F = B * N, H, w1.weight.shape[0]
            d.residual = 1 if op.residual else 0
            d.accumulate_grads = 1
            ln = op.ln if op.norm else None
            d.eps = ln.eps if ln is not None else 1e-6
            if training and op.DROPOUT_R > 0:
                d.rng_state = self.rng.data_ptr()
                d.p_mid = d.p_out = op.DROPOUT_R
                d.salt_mid, d.salt_out = (op._sites[0] << 32) | 1, (op._sites[1] << 32) | 1
            d.x, d.x16 = xs[i].data_ptr(), _lib.ptr(x16s[i])
            d.W1, d.b1, d.W2, d.b2 = w1.weight.data_ptr(), w1.bias.data_ptr(), w2.weight.data_ptr(), w2.bias.data_ptr()
            if bf:
                d.w16_1, d.w16_2 = self._managed(op, 'w1'), self._managed(op, 'w2')
            cd.params = [w1.weight, w1.bias, w2.weight, w2.bias]
            d.dW1, d.db1, d.dW2, d.db2 = (_grad_ptr(p) for p in cd.params)
            if ln is not None:
                d.ln_a, d.ln_b = ln.a_2.data_ptr(), ln.b_2.data_ptr()
                d.dln_a, d.dln_b = _grad_ptr(ln.a_2), _grad_ptr(ln.b_2)
                cd.params += [ln.a_2, ln.b_2]
            d.dout, d.dx = dxs[i + 1].data_ptr(), dxs[i].data_ptr()
            cd.fwd, cd.bwd = lib.mmnas_ffn_ln_fwd, lib.mmnas_ffn_ln_bwd
        elif isinstance(op, (M.SelfAtt, M.RelSelfAtt, M.GuidedAtt)):
            d = _lib.AttBlock()
            mh = op.mhatt
            if mh.HBASE != 64:
                raise NotImplementedError("only the '*_64' attention operators are implemented in CUDA")
            guided = isinstance(op, M.GuidedAtt)
            rel = isinstance(op, M.RelSelfAtt)
            cd.kind, cd.guided, cd.rel = 'att', guided, rel
            d.precision = 1 if bf else 0
            d.B, d.Nq, d.Nk, d.H, d.I = B, N, (Nx if guided else N), H, mh.HSIZE_INSIDE
            d.R = mh.linear_r.weight.shape[1] if rel else 0
            d.residual = 1 if op.residual else 0
            d.guided = 1 if guided else 0
            d.accumulate_grads = d.accumulate_geometry = d.accumulate_dkv = 1
            ln = op.ln if op.norm else None
            d.eps = ln.eps if ln is not None else 1e-6
            if training and op.DROPOUT_R > 0:
                d.rng_state = self.rng.data_ptr()
                d.p_att = d.p_out = op.DROPOUT_R
                d.salt_att, d.salt_out = (mh._sites[0] << 32) | 1, (mh._sites[1] << 32) | 1
            d.x, d.x16 = xs[i].data_ptr(), _lib.ptr(x16s[i])
            if guided:
                d.kv, d.kv16 = self.xs[-1].data_ptr(), _lib.ptr(self.x16s[-1])
                d.kmask = self.xmask.data_ptr()
                d.dkv = self.dxs[-1].data_ptr()
            else:
                d.kmask = (self.xmask if side == 'x' else self.ymask).data_ptr()
            q, k, v, m = mh.linear_q.weight, mh.linear_k.weight, mh.linear_v.weight, mh.linear_merge.weight
            d.Wq, d.Wk, d.Wv, d.Wm = q.data_ptr(), k.data_ptr(), v.data_ptr(), m.data_ptr()
            if bf:
                d.w16_a = self._managed(mh, 'q' if guided else 'vkq')
                if guided:
                    d.w16_b = self._managed(mh, 'vk')
                d.w16_m = self._managed(mh, 'm')
            cd.params = [v, k, q, m]
            d.dWq, d.dWk, d.dWv, d.dWm = _grad_ptr(q), _grad_ptr(k), _grad_ptr(v), _grad_ptr(m)
            if ln is not None:
                d.ln_a, d.ln_b = ln.a_2.data_ptr(), ln.b_2.data_ptr()
                d.dln_a, d.dln_b = _grad_ptr(ln.a_2), _grad_ptr(ln.b_2)
                cd.params += [ln.a_2, ln.b_2]
            if rel:
                wy, by = self.geo_params
                r = mh.linear_r
                d.g4, d.Wy, d.by = self.g4.data_ptr(), wy.data_ptr(), by.data_ptr()
                d.Wr, d.br = r.weight.data_ptr(), r.bias.data_ptr()
                d.dWy, d.dby, d.dWr, d.dbr = _grad_ptr(wy), _grad_ptr(by), _grad_ptr(r.weight), _grad_ptr(r.bias)
                cd.params += [r.weight, r.bias]
            d.dout, d.dx = dxs[i + 1].data_ptr(), dxs[i].data_ptr()
            cd.fwd, cd.bwd = ((lib.mmnas_rel_mha_ln_fwd, lib.mmnas_rel_mha_ln_bwd) if rel
                              else (lib.mmnas_mha_ln_fwd, lib.mmnas_mha_ln_bwd))
        else:
            raise NotImplementedError('executor: candidate %s' % type(op).__name__)
        cd.desc = d
        cd.ref = ctypes.byref(d)
        return cd

    @staticmethod
    def _managed(owner, name):
        ent = owner._w16.get(name)
        if ent is None or ent[0] != 'managed':
            raise RuntimeError('executor: bf16 weight shadows are not managed (build engine.WeightShadows first)')
        return ent[1].data_ptr()

    # ------------------------------------------------------------------------------------------------ execution
    def _call(self, fn, cd, stream, side):
        cd.desc.stream = stream
        cd.desc.side_stream = side
        rc = fn(cd.ref)
        if rc != 0:
            raise _lib.MMnasLibraryError('block call failed (%d): %s' % (rc, _lib.load().mmnas_last_error().decode()))

    def _cand_out(self, side, i, c):
        key = (side, i, c)
        buf = self.cand_out.get(key)
        if buf is None:
            ref = self.xs[0] if side == 'x' else self.ys[0]
            buf = self.cand_out[key] = torch.empty_like(ref)
        return buf

    def prepare(self, x_in, y_in):
        """(Re)build the static plans when the problem shape, device, precision arm or train/eval mode changed."""
        B, Nx, H = x_in.shape
        Ny = y_in.shape[1]
        dev = x_in.device
        bf = runtime.get_precision() == 'bf16'
        key = (B, Nx, Ny, H, dev, bf, self.net.training)
        if self.key != key:
            self._build(B, Nx, Ny, H, dev, bf)
            self.key = key

    def stage(self, x_in, y_in, x_mask, y_mask, g4):
        """Copy the stem's outputs into the static input buffers (device-side work only: capturable)."""
        self.xs[0].copy_(x_in)
        self.ys[0].copy_(y_in)
        self.xmask.copy_(x_mask.reshape(self.xmask.shape))
        self.ymask.copy_(y_mask.reshape(self.ymask.shape))
        self.g4.copy_(g4)
        if self.bf:
            K.cast_bf16(self.xs[0], self.x16s[0])
            K.cast_bf16(self.ys[0], self.y16s[0])

    def run_forward(self):
        """The 30 nodes over the staged inputs: the sampled candidate per node (MODE None) or every candidate and
        the gate-weighted sum (MODE 'full').  Results in xs[-1] / ys[-1]."""
        bf = self.bf
        full = MixedOp.MODE == 'full'
        self.full = full
        stream = _lib.stream()
        self.picks_x = [m.active_index[0] for m in self.nodes_x]
        self.picks_y = [m.active_index[0] for m in self.nodes_y]
        for side, cands, picks, bufs, b16s, gates in (('x', self.cands_x, self.picks_x, self.xs, self.x16s, self.gates_x),
                                                      ('y', self.cands_y, self.picks_y, self.ys, self.y16s, self.gates_y)):
            for i, row in enumerate(cands):
                if not full:
                    cd = row[picks[i]]
                    cd.desc.out, cd.desc.out16 = bufs[i + 1].data_ptr(), _lib.ptr(b16s[i + 1])
                    self._call(cd.fwd, cd, stream, None)
                else:
                    outs = []
                    for c, cd in enumerate(row):
                        o = self._cand_out(side, i, c)
                        cd.desc.out, cd.desc.out16 = o.data_ptr(), None
                        self._call(cd.fwd, cd, stream, None)
                        outs.append(o)
                    K.mixed_accum(outs, gates[i].data, bufs[i + 1])
                    if bf:
                        K.cast_bf16(bufs[i + 1], b16s[i + 1])

    def forward(self, x_in, y_in, x_mask, y_mask, g4):
        self.prepare(x_in, y_in)
        self.stage(x_in, y_in, x_mask, y_mask, g4)
        self.run_forward()
        return self.xs[-1].detach(), self.ys[-1].detach()

    def backward(self, dx_out, dy_out):
        if dx_out is None:
            self.dxs[-1].zero_()
        else:
            self.dxs[-1].copy_(dx_out)
        if dy_out is None:
            self.dys[-1].zero_()
        else:
            self.dys[-1].copy_(dy_out)
        return self.run_backward()

    def run_backward(self):
        """Backward of run_forward() from the output gradients staged in dxs[-1] / dys[-1]; input gradients in
        dxs[0] / dys[0], parameter gradients accumulated straight into the flat gradient buffer."""
        stream = _lib.stream()
        dev = self.xs[0].device
        side_stream = runtime.side_stream(dev).cuda_stream if (self.bf and runtime.overlap_wgrad) else None
        full = self.full
        listener = runtime.grad_listener
        used_rel = False
        for side, cands, picks, bufs, dbufs, gates, dcand in (
                ('y', self.cands_y, self.picks_y, self.ys, self.dys, self.gates_y, self.dcand_y),
                ('x', self.cands_x, self.picks_x, self.xs, self.dxs, self.gates_x, self.dcand_x)):
            for i in range(len(cands) - 1, -1, -1):
                row = cands[i]
                cd = row[picks[i]]
                if full:
                    # every alpha_gate gradient <o_k, dOut> of the node and the active candidate's gradient in one pass
                    outs = [self._cand_out(side, i, c) for c in range(len(row))]
                    d_outs = [dcand if c == picks[i] else None for c in range(len(row))]
                    K.mixed_alpha_dot(outs, gates[i].data, dbufs[i + 1], gates[i].grad, d_outs)
                    cd.desc.dout = dcand.data_ptr()
                else:
                    cd.desc.dout = dbufs[i + 1].data_ptr()
                self._call(cd.bwd, cd, stream, side_stream)
                used_rel = used_rel or cd.rel
                if listener is not None:
                    runtime.notify_grads(cd.params)
                    if full:
                        runtime.notify_grads((gates[i],))
            if side == 'y' and listener is not None and used_rel:
                runtime.notify_grads(self.geo_params)
        return self.dxs[0], self.dys[0]


class BackboneFn(Function):
    """The 30-node supernet backbone as one autograd node over a SearchExecutor."""

    @staticmethod
    def forward(ctx, x_in, y_in, x_mask, y_mask, g4, executor):
        _lib.require_cuda(x_in, y_in, g4)
        ctx.set_materialize_grads(False)
        ctx.executor = executor
        x_out, y_out = executor.forward(x_in.contiguous(), y_in.contiguous(), x_mask, y_mask, g4.contiguous())
        return x_out, y_out

    @staticmethod
    def backward(ctx, dx_out, dy_out):
        dx_in, dy_in = ctx.executor.backward(dx_out, dy_out)
        return dx_in, dy_in, None, None, None, None
