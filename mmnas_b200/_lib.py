"""ctypes binding of libmmnas_b200.so (the C ABI declared in include/mmnas_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, this raises.  PyTorch is
used for device memory and streams only; every argument crossing this boundary is a raw device
pointer, a size or a scalar."""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libmmnas_b200.so')
ABI_VERSION = 8

c_p, c_i, c_l, c_f, c_u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float, ctypes.c_ulonglong

# name -> argtypes, in header order
SIGNATURES = {
    'mmnas_gemm_f32': [c_i, c_i, c_i, c_p, c_l, c_l, c_p, c_l, c_l, c_p, c_l, c_p, c_i, c_i, c_p, c_l, c_f, c_p, c_u64,
                       c_f, c_p],
    'mmnas_gemm_bf16': [c_i, c_i, c_i, c_p, c_l, c_i, c_p, c_l, c_i, c_p, c_l, c_i, c_p, c_i, c_i, c_p, c_l, c_f, c_i,
                        c_p, c_u64, c_f, c_p],
    'mmnas_gemm_ln_bf16': [c_i, c_i, c_i, c_p, c_l, c_p, c_l, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_u64, c_f, c_p],
    'mmnas_attn_fwd': [c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_l, c_p, c_l, c_p, c_l, c_p, c_p, c_p, c_l, c_f, c_p, c_u64,
                       c_f, c_p],
    'mmnas_attn_bwd': [c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_l, c_p, c_l, c_p, c_l, c_p, c_p, c_p, c_l, c_p, c_l, c_p,
                       c_l, c_p, c_l, c_p, c_l, c_p, c_f, c_p, c_u64, c_f, c_p],
    'mmnas_relbias_fwd': [c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    'mmnas_relbias_bwd': [c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    'mmnas_ln_residual_fwd': [c_i, c_i, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p, c_p, c_u64, c_f, c_p],
    'mmnas_ln_residual_bwd': [c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_i, c_p, c_p, c_p, c_u64, c_f, c_p],
    'mmnas_mixed_accum': [c_i, c_p, c_p, c_p, c_l, c_p],
    'mmnas_mixed_alpha_dot': [c_i, c_p, c_p, c_p, c_p, c_p, c_l, c_p],
    'mmnas_cast_f32_to_bf16': [c_p, c_p, c_l, c_p],
    'mmnas_cast_multi': [c_p, c_i, c_p],
    'mmnas_colsum': [c_i, c_p, c_i, c_i, c_l, c_p, c_i, c_p],
    'mmnas_cast_rowmask': [c_p, c_p, c_p, c_i, c_i, c_p],
    'mmnas_sumsq_f32': [c_p, c_l, c_p, c_p, c_p],
    'mmnas_clip_adam': [c_p, c_i, c_p, c_p, c_p, c_f, c_f, c_f, c_f, c_p],
    'mmnas_rng_advance': [c_p, c_p],
    'mmnas_rowmask_bf16': [c_p, c_p, c_i, c_i, c_p],
    'mmnas_box_geometry': [c_p, c_p, c_p, c_i, c_i, c_p],
    'mmnas_lstm_workspace': [c_i, c_i, c_i, c_p],
    'mmnas_lstm_fwd': [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p],
    'mmnas_lstm_bwd': [c_i, c_i, c_i, c_p, c_p, c_p, c_p],
}



class AttBlock(ctypes.Structure):
    """Mirror of `mmnas_att_block` (include/mmnas_b200.h): descriptor of one SelfAtt / GuidedAtt / RelSelfAtt call."""
    _fields_ = ([(n, c_i) for n in ('precision', 'B', 'Nq', 'Nk', 'H', 'I', 'R', 'residual', 'guided', 'accumulate_grads',
                                    'accumulate_geometry', 'accumulate_dkv')] +
                [(n, c_f) for n in ('eps', 'p_att', 'p_out')] +
                [(n, c_u64) for n in ('salt_att', 'salt_out')] +
                [(n, c_p) for n in ('rng_state', 'x', 'x16', 'kv', 'kv16', 'kmask', 'Wq', 'Wk', 'Wv', 'Wm', 'w16_a', 'w16_b',
                                    'w16_m', 'ln_a', 'ln_b', 'rel', 'g4', 'Wy', 'by', 'Wr', 'br', 'out', 'out16', 'workspace',
                                    'dout', 'dx', 'dkv', 'dWq', 'dWk', 'dWv', 'dWm', 'dln_a', 'dln_b', 'dWy', 'dby', 'dWr',
                                    'dbr', 'drel', 'bwd_workspace', 'stream', 'side_stream')])


class FfnBlock(ctypes.Structure):
    """Mirror of `mmnas_ffn_block`: descriptor of one FeedForward call."""
    _fields_ = ([(n, c_i) for n in ('precision', 'M', 'H', 'F', 'residual', 'accumulate_grads')] +
                [(n, c_f) for n in ('eps', 'p_mid', 'p_out')] +
                [(n, c_u64) for n in ('salt_mid', 'salt_out')] +
                [(n, c_p) for n in ('rng_state', 'x', 'x16', 'W1', 'b1', 'W2', 'b2', 'w16_1', 'w16_2', 'ln_a', 'ln_b', 'out',
                                    'out16', 'workspace', 'dout', 'dx', 'dW1', 'db1', 'dW2', 'db2', 'dln_a', 'dln_b',
                                    'bwd_workspace', 'stream', 'side_stream')])


_pAtt, _pFfn, _pU64 = ctypes.POINTER(AttBlock), ctypes.POINTER(FfnBlock), ctypes.POINTER(c_u64)
SIGNATURES.update({
    'mmnas_att_block_workspace': [_pAtt, _pU64, _pU64],
    'mmnas_ffn_block_workspace': [_pFfn, _pU64, _pU64],
    'mmnas_mha_ln_fwd': [_pAtt], 'mmnas_mha_ln_bwd': [_pAtt],
    'mmnas_rel_mha_ln_fwd': [_pAtt], 'mmnas_rel_mha_ln_bwd': [_pAtt],
    'mmnas_ffn_ln_fwd': [_pFfn], 'mmnas_ffn_ln_bwd': [_pFfn],
})
_lib = None


class MMnasLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MMnasLibraryError(
            'libmmnas_b200.so is not built (%s). Run `python -m mmnas_b200.build`; there is no CPU or '
            'PyTorch fallback for the operator hot path.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.mmnas_abi_version.restype = c_i
    lib.mmnas_launch_count.restype = c_u64
    lib.mmnas_last_error.restype = ctypes.c_char_p
    if lib.mmnas_abi_version() != ABI_VERSION:
        raise MMnasLibraryError('libmmnas_b200.so ABI %d != binding ABI %d; rebuild' % (lib.mmnas_abi_version(), ABI_VERSION))
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_i
    for name, cls in (('mmnas_att_block_sizeof', AttBlock), ('mmnas_ffn_block_sizeof', FfnBlock)):
        getattr(lib, name).restype = c_i
        if getattr(lib, name)() != ctypes.sizeof(cls):
            raise MMnasLibraryError('%s() = %d but the ctypes mirror has %d bytes: header and binding disagree'
                                    % (name, getattr(lib, name)(), ctypes.sizeof(cls)))
    _lib = lib
    return lib


def workspace_bytes(desc):
    """(forward workspace bytes, backward scratch bytes) of a block descriptor."""
    fwd, bwd = c_u64(0), c_u64(0)
    name = 'mmnas_att_block_workspace' if isinstance(desc, AttBlock) else 'mmnas_ffn_block_workspace'
    rc = getattr(load(), name)(ctypes.byref(desc), ctypes.byref(fwd), ctypes.byref(bwd))
    if rc != 0:
        raise MMnasLibraryError('%s failed (%d): %s' % (name, rc, load().mmnas_last_error().decode()))
    return fwd.value, bwd.value


def call_block(name, desc):
    """One block-level foreign call."""
    lib = load()
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(ctypes.byref(desc))
        e1.record()
        _profile.append((name, desc, e0, e1))
    else:
        rc = getattr(lib, name)(ctypes.byref(desc))
    if rc != 0:
        raise MMnasLibraryError('%s failed (%d): %s' % (name, rc, lib.mmnas_last_error().decode()))


def launches():
    """Kernels launched by the library in this process so far (counted inside the library: bench.py's `gpu_launches`)."""
    return int(load().mmnas_launch_count())


_profile = None       # when a list: (name, args, start_event, end_event) per call — bench.py's live kernel timing


def call(name, *args):
    lib = load()
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        _profile.append((name, args, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise MMnasLibraryError('%s failed (%d): %s' % (name, rc, lib.mmnas_last_error().decode()))


def profile_begin():
    """Start recording a CUDA-event pair around every C-ABI call on the current stream."""
    global _profile
    _profile = []


def profile_end():
    """Stop recording; returns [(entry point, args, milliseconds)] after synchronising."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    return [(n, a, e0.elapsed_time(e1)) for n, a, e0, e1 in rec]


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_raw_device = getattr(torch._C, '_cuda_getDevice', None)


_tls = threading.local()


def set_stream_override(handle):
    """Route this thread's C-ABI launches to the given cudaStream_t (None = back to torch's current stream).  Used for
    the side-stream weight-gradient work of a block backward: cheaper than torch.cuda.stream() context switches on a
    path that only launches library kernels."""
    _tls.override = handle


def stream():
    """cudaStream_t of torch's current stream on the current device (raw query: this sits on every launch path)."""
    o = getattr(_tls, 'override', None)
    if o is not None:
        return o
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    """CUDA tensors only, and on the CURRENT device: the launch stream is resolved from the current device, so a
    tensor of another GPU would be launched on the wrong device / stream."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise MMnasLibraryError('mmnas_b200 operators run on CUDA tensors only (got a %s tensor); '
                                    'there is no CPU path.' % t.device)
        if _raw_device is not None and t.device.index != _raw_device():
            raise MMnasLibraryError('tensor lives on cuda:%d but the current device is cuda:%d: wrap the call in '
                                    'torch.cuda.device(tensor.device)' % (t.device.index, _raw_device()))


def ptr_array(tensors):
    """Host array of device pointers (NULL for None entries)."""
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
