"""bench.py — MMnas-VQA train throughput on B200 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                      the reference's own modules (baseline/_ref) on host cores

A step = one pass of the hot path over one synthetic batch: the train step of train_vqa.py:294-311 (forward of
Net_Full with arch mmnas_vqa — 12 encoder + 18 decoder blocks through the CUDA operators —, BCE-sum loss,
backward, gradient mean over ranks, clip_grad_norm_ 1.0, Adam) at B=64 per GPU, dropout 0.1, bf16 arm.
`value` / `e2e` / `roofline` describe that workload (BASELINE configs[1]).  The same JSON line carries a `workloads`
object with the other BASELINE configs measured at the run's N through the same engine: the supernet search step
(configs[2]: weight step, architecture step, the 4:1 mix of ALPHA_EVERY=5), the VGD step (configs[3]) and the ITM step
(configs[4]).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'MMnas-VQA train samples/s'
WORKLOAD = ('BASELINE configs[1]: MMnas-VQA train step, arch mmnas_vqa (12 enc + 18 dec blocks, H=512, 8 heads), '
            'batch 64 per GPU, synthetic 100x2048 region features + boxes, 14-token questions, vocab 20000, '
            '3129 answers, dropout 0.1, fwd+bwd+clip+Adam')
BATCH = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-workloads', action='store_true', help='skip the search / VGD / ITM workloads')
    ap.add_argument('--workload-steps', type=int, default=20)
    ap.add_argument('--profile-out', default=None, help='write the live per-kernel table (JSON) here')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return {'hbm': d['hbm_gbs'], 'tensor_burst': d['bf16_tflops'], 'tensor': d['bf16_tflops_sustained'], 'src': 'measured'}
    return {'hbm': 6650.0, 'tensor_burst': 1590.0, 'tensor': 1400.0, 'src': 'fallback'}


# ------------------------------------------------------------------------------------------ CPU reference arm
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, 'mmnas', 'model', 'full_vqa.py'))


def cpu_step_fn(batch, threads=None):
    """The reference's train step on the host cores.  With baseline/_ref present (the byte-identical copy of the
    reference's model package, scripts/install_reference.py) this is the UNMODIFIED reference: its full_vqa.Net_Full,
    its operators, its WarmupOptimizer, driven by the step body of train_vqa.py:294-311 — kind 'reference'.  Without
    it, the oracle port (oracle/mmnas_oracle.py) — kind 'port'."""
    import torch
    from mmnas_b200 import genotypes
    from mmnas_b200.data.synthetic import Cfg, SynthSpec, make_batch, init_dict
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(888)
    spec = SynthSpec(batch=batch)
    cfg = Cfg(genotype=genotypes.shipped('mmnas_vqa'))
    inputs, target = make_batch(spec)
    if reference_available():
        sys.path.insert(1, REF_DIR)
        from mmnas.model.full_vqa import Net_Full          # the reference's own net, operators and optimizer wrapper
        from mmnas.utils.optimizer import WarmupOptimizer
        import mmnas.model.modules as ref_modules
        assert ref_modules.__file__.startswith(REF_DIR), 'reference arm must run the reference operators'
        net = Net_Full(cfg, init_dict(spec)).train()
        optim = WarmupOptimizer(cfg.NET_LR_BASE, torch.optim.Adam(net.parameters(), lr=0, betas=cfg.OPT_BETAS,
                                                                  eps=cfg.OPT_EPS, weight_decay=0), 10 ** 6, warmup=True)
        loss_fn = torch.nn.BCEWithLogitsLoss(reduction='sum')

        def step():                                        # train_vqa.py:294-311
            optim.zero_grad()
            pred = net(inputs)
            loss = loss_fn(pred, target)
            loss += 0 * sum(p.sum() for p in net.parameters())
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), cfg.NET_GRAD_CLIP)
            optim.step()
            return float(loss.detach())
        return step, 'reference'
    from oracle import mmnas_oracle as O
    from mmnas_b200.model.nets import Net_Full
    net = Net_Full(cfg, init_dict(spec))          # parameter container only; the arithmetic below is the oracle's
    P = O.leaf_params(net.state_dict(), torch.float32)
    params = [p for p in P.values() if p.requires_grad]
    state = {}
    counter = [0]

    def step():
        counter[0] += 1
        for p in params:
            p.grad = None
        loss, _ = O.train_step_vqa(P, inputs, target, cfg.GENOTYPE, p=cfg.DROPOUT_R, training=True)
        O.clip_and_adam(params, state, counter[0], lr=cfg.NET_LR_BASE / 4, max_norm=1.0)
        return float(loss)
    return step, 'port'


def cpu_baseline(seconds=12.0, batch=BATCH):
    import torch
    cores = os.cpu_count() or 1
    step, kind = cpu_step_fn(batch, cores)
    step()                                         # warm-up
    t0, n = time.perf_counter(), 0
    while n < 1 or (time.perf_counter() - t0 < seconds and n < 5):
        step()
        n += 1
    dt = (time.perf_counter() - t0) / n
    what = ('the unmodified reference (full_vqa.Net_Full + WarmupOptimizer from baseline/_ref)' if kind == 'reference'
            else 'the oracle port')
    return {'value': batch / dt, 'unit': 'samples/s', 'cores': torch.get_num_threads(), 'kind': kind,
            'sample': '%d train steps of %s, torch CPU fp32, batch %d, %.2f s/step, after 1 warm-up' % (n, what, batch, dt)}


def gpu_eager_reference(dev, steps=5, batch=BATCH):
    """The bar BASELINE.md names: the UNMODIFIED reference (baseline/_ref: full_vqa.Net_Full, its PyTorch operators, its
    WarmupOptimizer) running eagerly on the SAME GPU, float32 (TF32 off, as the reference leaves it), step body of
    train_vqa.py:294-311, B=64, dropout 0.1, device-resident batch, CUDA events.  None when baseline/_ref is absent."""
    if not reference_available():
        return None
    import torch
    from mmnas_b200 import genotypes
    from mmnas_b200.data.synthetic import Cfg, SynthSpec, make_batch, init_dict
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = True           # torch defaults: what the reference runs with
    try:
        sys.path.insert(1, REF_DIR)
        from mmnas.model.full_vqa import Net_Full
        from mmnas.utils.optimizer import WarmupOptimizer
        torch.manual_seed(888)
        spec = SynthSpec(batch=batch)
        cfg = Cfg(genotype=genotypes.shipped('mmnas_vqa'))
        inputs, target = make_batch(spec)
        net = Net_Full(cfg, init_dict(spec)).to(dev).train()
        optim = WarmupOptimizer(cfg.NET_LR_BASE, torch.optim.Adam(net.parameters(), lr=0, betas=cfg.OPT_BETAS, eps=cfg.OPT_EPS),
                                10 ** 6, warmup=True)
        loss_fn = torch.nn.BCEWithLogitsLoss(reduction='sum')
        inputs, target = tuple(t.to(dev) for t in inputs), target.to(dev)

        def step():
            optim.zero_grad()
            loss = loss_fn(net(inputs), target)
            loss += 0 * sum(p.sum() for p in net.parameters())
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), cfg.NET_GRAD_CLIP)
            optim.step()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del net, optim
        torch.cuda.empty_cache()
        return {'value': batch / ms * 1e3, 'unit': 'samples/s', 'ms_per_step': ms,
                'how': 'the unmodified reference (baseline/_ref) eager on this GPU: full_vqa.Net_Full + its PyTorch operators + '
                       'WarmupOptimizer, fp32 (matmul TF32 off), batch %d, dropout 0.1, device-resident batch, %d steps after 3 '
                       'warm-up, CUDA events' % (batch, steps)}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    batch = BATCH                                  # the full configs[1] batch: same config as the B200 arm
    step, kind = cpu_step_fn(batch, os.cpu_count())
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    val = batch / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'samples/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': batch,
                       'note': ('reference CPU path: the UNMODIFIED reference modules (baseline/_ref: full_vqa.Net_Full, '
                                'modules.py operators, WarmupOptimizer) through the step body of train_vqa.py:294-311, '
                                'all host cores, batch 64' if kind == 'reference' else
                                'reference CPU path = oracle port (baseline/_ref not present)')},
            'cpu_baseline': {'value': val, 'unit': 'samples/s', 'cores': torch.get_num_threads(), 'kind': kind,
                             'sample': '%d steps at batch %d' % (args.steps, batch)},
            'e2e': {'value': val, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '50'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_max': max(power)}


def kernel_table(records):
    """Aggregate the live CUDA-event records per kernel family with algorithmic FLOPs / bytes."""
    fam = {}
    for name, a, ms in records:
        if name.endswith('_workspace') or name.endswith('_sizeof'):
            continue                    # size queries: foreign calls, not kernels
        flops = byts = 0.0
        key = name.replace('mmnas_', '')
        if name == 'mmnas_gemm_bf16':
            M, N, K_ = a[0], a[1], a[2]
            flops = 2.0 * M * N * K_
            byts = 2.0 * (M * K_ + N * K_) + (2.0 if a[11] else 4.0) * M * N
            key = 'gemm_bf16_tc'        # one kernel template; MMNAS_PROFILE_SHAPES=1 splits it by layout and shape
            if os.environ.get('MMNAS_PROFILE_SHAPES'):
                key = 'gemm_bf16[%s%s] %dx%dx%d%s' % ('mn' if a[5] else 'k', 'mn' if a[8] else 'k', M, N, K_,
                                                     ' sk%d' % a[18] if a[18] > 1 else '')
        elif name == 'mmnas_gemm_f32':
            flops = 2.0 * a[0] * a[1] * a[2]
            byts = 4.0 * (a[0] * a[2] + a[1] * a[2] + a[0] * a[1])
        elif name in ('mmnas_attn_fwd', 'mmnas_attn_bwd'):
            B, h, Nq, Nk = a[1], a[2], a[3], a[4]
            flops = (4.0 if name.endswith('fwd') else 10.0) * B * h * Nq * Nk * 64
            es = 4.0 if a[0] == 0 else 2.0
            byts = es * B * h * 64 * (2 * Nq + 2 * Nk) * (1 if name.endswith('fwd') else 2)
        elif name in ('mmnas_ln_residual_fwd', 'mmnas_ln_residual_bwd'):
            byts = 4.0 * a[0] * a[1] * 3
        elif name in ('mmnas_relbias_fwd', 'mmnas_relbias_bwd'):
            B, N, h = a[1], a[2], a[3]             # (mode, B, N, heads, R, rel, g4, ...)
            pairs = float(B) * N * N
            geometry = a[6] is not None            # g4 given: the 4 -> 64 layer is recomputed in the kernel
            flops = pairs * 2 * ((64 * 4 if geometry else 0) + 64 * h) * (1 if name.endswith('fwd') else 3)
            byts = pairs * ((16 if geometry else 256) + 4 * h) * (1 if name.endswith('fwd') else 2)
        elif name == 'mmnas_cast_f32_to_bf16':
            byts = 6.0 * a[2]
        elif name == 'mmnas_colsum':
            byts = (4.0 if a[0] == 0 else 2.0) * a[2] * a[3]
        f = fam.setdefault(key, {'launches': 0, 'ms': 0.0, 'flops': 0.0, 'bytes': 0.0})
        f['launches'] += 1; f['ms'] += ms; f['flops'] += flops; f['bytes'] += byts
    return fam


def run_workloads(args, world, rank, dev, barrier):
    """BASELINE configs[2..4] at the run's N: device-resident synthetic batches (as `value`), CUDA events, max over
    ranks, B=64 per GPU, bf16 arm, dropout 0.1, data-parallel gradient mean where N > 1."""
    import torch
    import torch.distributed as dist
    import mmnas_b200
    from mmnas_b200 import _lib, genotypes
    from mmnas_b200.data.synthetic import Cfg, make_batch, init_dict, spec_for
    from mmnas_b200.engine import TrainStep, SearchStep, LOSSES, tree_map
    from mmnas_b200.model.nets import Net_Full, Net_Search
    steps = max(3, min(args.steps, args.workload_steps))

    def timed(fn, n):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launches()
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / n, float(out), (_lib.launches() - l0) // n

    def to_dev(batch):
        return tree_map(lambda t: t.to(dev), batch)

    out = {}
    # ---- configs[2]: supernet search step (search_vqa.py:278-337), H=256, 4 heads, 12 + 18 MixedOp nodes
    torch.manual_seed(888)
    spec = spec_for('vqa', batch=BATCH)
    cfg = Cfg(mode='search')
    net = Net_Search(cfg, init_dict(spec)).to(dev).train()
    mmnas_b200.manual_seed(888 + rank, dev)
    train_b, eval_b = to_dev(make_batch(spec, seed=2000 + 17 * rank)), to_dev(make_batch(spec, seed=3000 + 17 * rank))
    sstep = SearchStep(net, lr_base=cfg.NET_LR_BASE, epoch_steps=10 ** 6, alpha_lr=cfg.ALPHA_LR_BASE,
                       alpha_betas=cfg.ALPHA_OPT_BETAS, mode=cfg.ALPHA_BINARY_MODE)
    tw, lw, nw = timed(lambda: sstep.weight_step(*train_b), steps)
    ta, la, na = timed(lambda: sstep.arch_step(*eval_b), steps)
    nw, na = nw + sstep.replayed_launches(False), na + sstep.replayed_launches(True)     # eager block calls + graph segments
    every = cfg.ALPHA_EVERY
    out['search_vqa'] = {
        'config': 'BASELINE configs[2]: MMnas-VQA supernet search step, H=256, 4 heads, MixedOp over SA/FFN (12 enc nodes) '
                  'and SA/RSA/GA/FFN (18 dec nodes), batch 64 per GPU, dropout 0.1; the sampled backbone is launched eagerly '
                  '(one foreign call per block), stem / heads + loss / stem backward / clip + Adam replay from four CUDA '
                  'graphs; seed-888 sampling with the reference\'s RNG consumption, drawn one step ahead',
        'weight_step': {'ms_per_step': tw, 'samples_per_s': BATCH * world / (tw / 1e3), 'launches_per_step': nw,
                        'final_loss': lw},
        'arch_step': {'ms_per_step': ta, 'samples_per_s': BATCH * world / (ta / 1e3), 'launches_per_step': na,
                      'final_loss': la, 'mode': cfg.ALPHA_BINARY_MODE},
        'mix': {'samples_per_s': every * BATCH * world / ((every * tw + ta) / 1e3), 'unit': 'train samples/s',
                'how': 'ALPHA_EVERY=%d (search_vqa.py:150,305): %d weight steps + 1 architecture step per %d iterations'
                       % (every, every, every)},
        'steps': steps}
    del sstep, net, train_b, eval_b
    torch.cuda.empty_cache()

    # ---- configs[3] / configs[4]: VGD and ITM train steps (train_vgd.py:317-341, train_itm.py:384-397)
    for task, arch, lr, label in (
            ('vgd', 'mmnas_vgd', 0.00014, 'BASELINE configs[3]: MMnas-VGD train step, arch mmnas_vgd (RSA-heavy: 5 RSA + 11 GA '
             'blocks), H=512, 100 valid regions + boxes, 15-token queries, KLD + 0.5 SmoothL1 loss, batch 64 per GPU'),
            ('itm', 'mmnas_itm', 0.00015, 'BASELINE configs[4]: MMnas-ITM train step, arch mmnas_itm, H=512, 36 regions, 50-token '
             'captions, batch 64 per GPU = 64 positive pairs + 64 negative-caption + 64 negative-image pairs: the '
             'reference\'s three forwards per step run as one stacked batch of 192 pairs; BCE_Loss (positive term twice)')):
        torch.manual_seed(888)
        spec = spec_for(task, batch=BATCH)
        cfg = Cfg(genotype=genotypes.shipped(arch), SCORES_LOSS='kld')
        net = Net_Full(cfg, init_dict(spec), task=task).to(dev).train()
        mmnas_b200.manual_seed(888 + rank, dev)
        inputs, target = to_dev(make_batch(spec, seed=4000 + 17 * rank))
        use_graph = not args.no_graph and (world == 1 or os.environ.get('MMNAS_DP_GRAPH', '1') == '1')
        tstep = TrainStep(net, lr_base=lr, epoch_steps=10 ** 6, use_graph=use_graph, loss_fn=LOSSES[task])
        ms, loss, _ = timed(lambda: tstep(inputs, target), steps)
        out['train_' + task] = {'config': label, 'ms_per_step': ms, 'samples_per_s': BATCH * world / (ms / 1e3),
                                'pairs_forwarded_per_s': (3 if task == 'itm' else 1) * BATCH * world / (ms / 1e3),
                                'final_loss': loss, 'cuda_graph': use_graph, 'steps': steps}
        del tstep, net, inputs, target
        torch.cuda.empty_cache()
    out['note'] = ('same engine, timing rule and per-GPU batch as `value`; n_gpus = %d, gradient mean over ranks '
                   'inside every step when n_gpus > 1' % world)
    return out


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import mmnas_b200
    from mmnas_b200 import _lib, genotypes
    from mmnas_b200.data.synthetic import Cfg, SynthSpec, make_batch, init_dict, compact
    from mmnas_b200.engine import TrainStep, Prefetcher, ScalarLog
    from mmnas_b200.model.nets import Net_Full

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL_DEBUG is left as the caller set it: main() re-routed fd 1 to stderr, so NCCL's INFO lines cannot mix
        # with the single JSON line on stdout
        dist.init_process_group('nccl', device_id=dev)
    mmnas_b200.set_precision(args.precision)
    _lib.load()

    torch.manual_seed(888)                        # identical init on every rank (DDP broadcast equivalent)
    spec = SynthSpec(batch=BATCH)
    cfg = Cfg(genotype=genotypes.shipped('mmnas_vqa'))
    net = Net_Full(cfg, init_dict(spec)).to(dev).train()
    mmnas_b200.manual_seed(888 + rank, dev)       # per-rank dropout masks
    # compact loader format: bf16 region features + raw boxes; the [B,100,100,4] log-geometry the
    # reference's DataLoader builds per sample on the CPU is built on the device inside the step (mmnas_box_geometry)
    host_batches = [compact(make_batch(spec, seed=1000 + 17 * rank + i)) for i in range(2)]
    inputs, target = host_batches[0]
    dev_in, dev_tgt = tuple(t.to(dev) for t in inputs), target.to(dev)
    # the whole step (incl. the bucketed NCCL all-reduces launched from the backward hooks) is one CUDA graph;
    # MMNAS_DP_GRAPH=0 falls back to eager launches for N > 1
    use_graph = not args.no_graph and (world == 1 or os.environ.get('MMNAS_DP_GRAPH', '1') == '1')
    step = TrainStep(net, lr_base=cfg.NET_LR_BASE, epoch_steps=10 ** 6, use_graph=use_graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    for _ in range(max(3, args.warmup)):
        step(dev_in, dev_tgt)
    barrier()
    launches_before = _lib.launches()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(dev_in, dev_tgt)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches_eager = _lib.launches() - launches_before
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = BATCH * world / (ms_step / 1e3)
    loss_val = float(loss)

    # ---- end-to-end through the public step API with host buffers (`e2e`)
    pre = Prefetcher(host_batches, dev)
    for _ in range(3):
        b_in, b_t = pre.next()
        float(step(b_in, b_t))
    barrier()
    log = ScalarLog(dev)
    losses = []
    e0.record()
    for _ in range(args.steps):
        b_in, b_t = pre.next()
        log.push(step(b_in, b_t))                 # device->host copy of this step's loss, enqueued behind the step ...
        if len(log) > 1:
            losses.append(log.pop())              # ... and read on the host while the next step is already queued
    while len(log):
        losses.append(log.pop())
    last = losses[-1]
    e1.record()
    barrier()
    assert len(losses) == args.steps and all(v == v for v in losses)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    e2e = {'value': BATCH * world / (e2e_ms / 1e3), 'unit': 'samples/s', 'ms_per_step': e2e_ms,
           'h2d_bytes_per_step': pre.bytes_per_batch, 'd2h_bytes_per_step': 4,
           'how': 'pinned host batch (bf16 region features, raw boxes, question tokens, answer scores) -> device on a copy '
                  'stream one step ahead, TrainStep(...) incl. the box-geometry kernel, every step\'s loss copied to pinned '
                  'host memory behind the step and read by the host one step later (engine.ScalarLog), all inside the '
                  'timed region'}

    # ---- live per-kernel timing (CUDA events around every C-ABI call, eager, same workload) -> roofline
    prof_step = step if not use_graph else TrainStep(net, lr_base=cfg.NET_LR_BASE, epoch_steps=10 ** 6, use_graph=False)
    from mmnas_b200 import runtime as _rt
    prof_step(dev_in, dev_tgt)
    l0 = _lib.launches()
    prof_step(dev_in, dev_tgt)                     # the step as it runs in the timed region (block-level calls, fused tails)
    launches_per_step = _lib.launches() - l0
    _rt.overlap_wgrad = False       # instrumented pass: one kernel at a time on one stream, so each event pair times its kernel alone
    _rt.compose_in_python = True    # ... and one foreign call per kernel (the primitive entry points) instead of one per block
    prof_step(dev_in, dev_tgt)
    torch.cuda.synchronize()
    _lib.profile_begin()
    n_prof = 3
    for _ in range(n_prof):
        # park the GPU behind a ~40 ms spin so the whole step is queued before it runs: the event pairs then
        # bracket device time only (in eager mode an idle GPU would otherwise charge host launch latency to them)
        torch.cuda._sleep(int(8e7))
        prof_step(dev_in, dev_tgt)
        torch.cuda.synchronize()
    fam = kernel_table(_lib.profile_end())
    _rt.overlap_wgrad = True
    _rt.compose_in_python = False
    pk = peaks()
    tot_ms = sum(f['ms'] for f in fam.values())
    top_name, top = max(fam.items(), key=lambda kv: kv[1]['ms'])
    if top['flops'] / max(top['bytes'], 1.0) > 200:
        bound, ach, peak, unit = 'tensor', top['flops'] / (top['ms'] * 1e-3) / 1e12, pk['tensor'], 'TFLOP/s'
    else:
        bound, ach, peak, unit = 'hbm', top['bytes'] / (top['ms'] * 1e-3) / 1e9, pk['hbm'], 'GB/s'
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        if top_name in tr:
            traffic = tr[top_name]['dram_bytes_per_launch']
    except (OSError, ValueError, KeyError):
        pass
    roofline = {'kernel': top_name, 'bound': bound, 'achieved': ach, 'peak': peak, 'unit': unit, 'frac': ach / peak,
                'traffic': traffic, 'traffic_source': 'static: profiles/ncu_traffic.json (one ncu --set full capture of this kernel family; dram bytes per launch), not measured in this run', 'peak_source': pk['src'] + (' (sustained)' if bound == 'tensor' else ''),
                'avg_launch_ms': top['ms'] / top['launches'], 'launches_per_step': top['launches'] // n_prof,
                'algorithmic_per_launch': (top['flops'] if bound == 'tensor' else top['bytes']) / top['launches'],
                'share_of_kernel_time': top['ms'] / tot_ms,
                'how': 'CUDA events around every C-ABI launch on the launching stream, %d eager steps with the GPU parked '
                       'behind a spin kernel and the wgrad side stream disabled (kernels serialised); achieved = '
                       'algorithmic FLOPs (2MNK, padding excluded) of all launches / their summed durations' % n_prof}
    if args.profile_out and rank == 0:
        rows = {k: dict(v, ms_per_step=v['ms'] / n_prof, share=v['ms'] / tot_ms,
                        tflops=v['flops'] / max(v['ms'], 1e-9) / 1e9, gbs=v['bytes'] / max(v['ms'], 1e-9) / 1e6)
                for k, v in fam.items()}
        json.dump({'kernels': rows, 'kernel_ms_per_step': tot_ms / n_prof, 'step_ms': ms_step}, open(args.profile_out, 'w'),
                  indent=1)

    # ---- the other BASELINE configs at this N, through the same engine (`workloads`)
    workloads = None
    if not args.no_workloads:
        del step, prof_step, pre
        torch.cuda.empty_cache()
        workloads = run_workloads(args, world, rank, dev, barrier)

    cpu = gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gpu_ref = gpu_eager_reference(dev)
        cpu = cpu_baseline()

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': max(3, args.warmup), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'global_batch': BATCH * world, 'parallelism': 'dp%d' % world,
                           'cuda_graph': use_graph,
                           'input_format': 'compact: region features bf16 [B,100,2048], raw boxes [B,100,4] (pairwise '
                                           'log-geometry built on the device each step), tokens int64, answer scores fp32',
                           'l2': 'no flush: one step streams >1 GB of weights, activations and optimizer state, '
                                 'far more than the 126 MB L2',
                           'final_loss': loss_val, 'e2e_final_loss': last},
                'clocks': clocks, 'e2e': e2e,
                'gpu_launches': launches_per_step * args.steps,
                'gpu_launches_note': '%d kernels of libmmnas_b200 per step (counted inside the library over one eager step; %s)' %
                                     (launches_per_step, 'replayed from the captured CUDA graph in the timed region'
                                      if use_graph else 'launched eagerly'),
                'roofline': roofline}
        if workloads:
            line['workloads'] = workloads
        if cpu:
            line['cpu_baseline'] = cpu
        if gpu_ref:
            line['reference_gpu_eager'] = gpu_ref
        emit(line)
    if world > 1:
        # a captured graph that contains NCCL kernels keeps the communicator busy at teardown; all results are
        # out, so leave without the (blocking) communicator destruction
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, warnings from
    libraries writing to fd 1) was re-routed to stderr in main()."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
