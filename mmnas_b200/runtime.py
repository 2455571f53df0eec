"""Process-wide runtime state of the operator path: precision mode and the dropout RNG state."""
import itertools
import os
import threading

import torch

from . import kernels

_PRECISION = 'bf16'          # 'fp32' (FFMA kernels, 1e-5 parity arm) | 'bf16' (tcgen05 arm, 2e-2)
_rng_states = {}
shadows_fresh = False      # True while an engine step guarantees that the managed bf16 weight shadows are current
direct_grads = False       # True while an engine step wants block backwards to accumulate straight into p.grad
overlap_wgrad = os.environ.get('MMNAS_OVERLAP_WGRAD', '1') != '0'   # weight-gradient GEMMs on a side stream, concurrently with the dgrad / attention chain (env: ablation only)
native_lstm = os.environ.get('MMNAS_NATIVE_LSTM', '1') != '0'      # persistent LSTM kernels in the bf16 arm (csrc/lstm.cu)
compose_in_python = os.environ.get('MMNAS_COMPOSE_PY', '0') == '1'   # blocks composed from the primitive entry points in Python (one foreign call per kernel) instead of the block-level C calls: bench.py's per-kernel timing pass and the cross-check tests
grad_listener = None       # callable(param): the data-parallel reducer's notification for directly written grads
_salt_counter = itertools.count(1)
_lock = threading.Lock()


def set_precision(mode):
    global _PRECISION
    if mode not in ('fp32', 'bf16'):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _PRECISION = mode
    # The callers' stem / heads still run on torch: keep them true fp32 in the parity arm (cuDNN's LSTM would
    # otherwise use TF32, ~1e-3), and let them use TF32 tensor cores in the bf16 arm.
    torch.backends.cudnn.allow_tf32 = mode == 'bf16'
    torch.backends.cuda.matmul.allow_tf32 = mode == 'bf16'


def get_precision():
    return _PRECISION


class precision:
    """Context manager: `with mmnas_b200.precision('fp32'): ...`"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_precision()
        set_precision(self.mode)

    def __exit__(self, *a):
        set_precision(self.prev)


def new_site():
    """A process-unique id for one dropout site (one per module instance and per site inside it)."""
    with _lock:
        return next(_salt_counter)


def rng_state(device):
    """Device tensor {seed, step} (int64 bit patterns) driving every dropout site on `device`."""
    key = torch.device(device)
    if key.index is None:
        key = torch.device('cuda', torch.cuda.current_device())
    st = _rng_states.get(key)
    if st is None:
        seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        st = torch.tensor([seed, 0], dtype=torch.int64, device=key)
        _rng_states[key] = st
    return st


def manual_seed(seed, device=None):
    """Re-seed the dropout state of `device` (default: current CUDA device); the step counter restarts at 0."""
    st = rng_state(device if device is not None else torch.device('cuda', torch.cuda.current_device()))
    st.copy_(torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64))


# Host-side twin of the step counter.  Dropout sites number their calls WITHIN a step: a module that runs several times
# per step (the ITM forwards of the per-module path) gets a fresh mask per call; across steps the device counter supplies
# the fresh stream.  A captured step (call index baked in at capture) and an eager step therefore draw the same masks.
epoch = 0


def advance(device=None):
    """Bump the step counter: call once per training step (captured fine inside a CUDA graph)."""
    global epoch
    epoch += 1
    st = rng_state(device if device is not None else torch.device('cuda', torch.cuda.current_device()))
    kernels.rng_advance(st)


def notify_grads(params):
    """Tell the data-parallel reducer that these parameters' gradients were written in place (no autograd hook fires
    for a gradient a backward returns as None)."""
    if grad_listener is not None:
        for p in params:
            if p is not None:
                grad_listener(p)


_side_streams = {}


def side_stream(device):
    """The per-device side stream for work that is independent of the main backward chain (weight gradients)."""
    key = torch.device(device)
    st = _side_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=key)
        _side_streams[key] = st
    return st
