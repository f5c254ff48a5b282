"""ctypes binding of the C-ABI library ``libmmlf_b200.so`` (declared in include/mmlf_b200.h).

The library is built in-tree by ``mmlf_b200/csrc/Makefile`` (``__graft_entry__.build()``).  There is no CPU
fallback: if the library is missing, or a kernel is called without an sm_100 device, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmmlf_b200.so')

c_p = C.c_void_p
c_i = C.c_int
c_i64 = C.c_int64
c_d = C.c_double
c_f = C.c_float


class ConvArgs(C.Structure):
    """Mirror of ``mmlf_conv_args`` (include/mmlf_b200.h)."""
    _fields_ = [('in_', c_p), ('ld_in', c_i), ('cin_pad', c_i), ('wpack', c_p), ('n_pad', c_i),
                ('B', c_i), ('H', c_i), ('W', c_i), ('type', c_i),
                ('bias', c_p), ('scale', c_p), ('shift', c_p), ('relu', c_i),
                ('gate_bits', c_p), ('relu_bits', c_p), ('ld_bits', c_i),
                ('out', c_p), ('ld_out', c_i), ('out_mode', c_i), ('n_real', c_i),
                ('out2', c_p), ('ld_out2', c_i), ('col_sums', c_p),
                ('ab_dtype', c_i), ('out_dtype', c_i), ('out2_dtype', c_i),
                ('split_in', c_i), ('split_out', c_i),
                ('bn_z', c_p), ('ld_z', c_i), ('bn_z_dtype', c_i), ('bn_scale', c_p), ('bn_shift', c_p), ('bn_mean', c_p)]


_PROTOS = {
    'mmlf_last_error': (C.c_char_p, []),
    'mmlf_abi_version': (c_i, []),
    'mmlf_check_device': (c_i, []),
    'mmlf_lf_extract_u8': (c_i, [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    'mmlf_lf_shift': (c_i, [c_p] * 8 + [c_i, c_i, c_i, c_i, c_d, c_p]),
    'mmlf_texture_mask': (c_i, [c_p, c_i, c_i, c_i, c_i, c_d, c_p, c_p, c_p]),
    'mmlf_lmm_to_discrete': (c_i, [c_p, c_p, c_i, c_i64, c_i64, c_i, c_d, c_d, c_p, c_p]),
    'mmlf_kl_divergence': (c_i, [c_p, c_p, c_i, c_i64, c_i64, c_p, c_p, c_p, c_p]),
    'mmlf_nll_discrete': (c_i, [c_p, c_p, c_i, c_i64, c_i64, c_p, c_p, c_p, c_p]),
    'mmlf_augment_fill': (c_i, [c_p, c_i, c_d, c_i]),
    'mmlf_augment_patches': (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_p,
                                   c_p, c_p, c_p]),
    'mmlf_augment_contrast': (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p]),
    'mmlf_shift_taps': (c_i, [c_d, c_i, C.POINTER(c_f), C.POINTER(c_f), C.POINTER(c_i), C.POINTER(c_i)]),
    'mmlf_pack_views': (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p]),
    'mmlf_pack_views_split': (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p]),
    'mmlf_shift_pack': (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_d, c_p, c_i, c_i, c_p]),
    'mmlf_pack_stacks': (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_d, c_p]),
    'mmlf_pack_conv_weight': (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_p]),
    'mmlf_pack_conv_weight_split': (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_f, c_p]),
    'mmlf_pack_conv_weights_batch': (c_i, [c_p, c_i, c_i64, c_p]),
    'mmlf_unpack_conv_wgrad': (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p]),
    'mmlf_conv2x2': (c_i, [C.POINTER(ConvArgs), c_p]),
    'mmlf_conv2x2_simt': (c_i, [C.POINTER(ConvArgs), c_p]),
    'mmlf_conv2x2_wgrad_workspace': (c_i64, [c_i, c_i]),
    'mmlf_conv2x2_wgrad': (c_i, [c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    'mmlf_conv2x2_wgrad_canonical': (c_i, [c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p,
                                           c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p]),
    'mmlf_convert16': (c_i, [c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_i64, c_p]),
    'mmlf_colsum16': (c_i, [c_p, c_i, c_i, c_i64, c_i, c_p, c_i, c_p]),
    'mmlf_bn_stats': (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    'mmlf_bn_finalize': (c_i, [c_p, c_i, c_i, c_i64, c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_p, c_p]),
    'mmlf_bn_fold_eval': (c_i, [c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_p]),
    'mmlf_bn_apply_relu': (c_i, [c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p]),
    'mmlf_bn_bwd_reduce': (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    'mmlf_bn_bwd_apply': (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i, c_i, c_i, c_i, c_i,
                                c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p]),
    'mmlf_relu_bwd': (c_i, [c_p, c_i, c_p, c_i, c_i, c_i64, c_i, c_i, c_p, c_i, c_p]),
    'mmlf_head_small': (c_i, [c_p, c_i, c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    'mmlf_head_small_bwd': (c_i, [c_p, c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_p]),
    'mmlf_upr_posterior': (c_i, [c_p, c_p, c_p, c_i, c_i64, c_i64, c_p, c_p]),
    'mmlf_dpp_head': (c_i, [c_p, c_p, c_p, c_i, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p]),
    'mmlf_reg_to_class': (c_i, [c_p, c_p, c_i, c_d, c_i64, c_i64, c_p, c_p]),
    'mmlf_mpi_to_weights': (c_i, [c_p, c_i, c_p, c_i, c_d, c_i64, c_i64, c_p, c_p]),
    'mmlf_loss_prepass': (c_i, [c_p, c_p, c_p, c_i, c_i64, c_i64, c_p, c_p]),
    'mmlf_loss_regression': (c_i, [c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_d, c_i64, c_i64, c_p, c_p, c_p, c_i64, c_p]),
    'mmlf_loss_cross_entropy': (c_i, [c_p, c_p, c_p, c_p, c_d, c_i, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p]),
    'mmlf_ese_reduce': (c_i, [c_p, c_p, c_p, c_i, c_i64, c_i64, c_p, c_p, c_p, c_p]),
    'mmlf_adam_step': (c_i, [c_p, c_p, c_p, c_p, c_i64, c_d, c_d, c_d, c_d, c_i64, c_p]),
    'mmlf_adam_step_dev': (c_i, [c_p, c_p, c_p, c_p, c_i64, c_p, c_d, c_d, c_d, c_p]),
    'mmlf_zero': (c_i, [c_p, c_i64, c_p]),
    'mmlf_vec_jobs': (c_i, [c_p, c_i, c_p]),
    'mmlf_loss_finish': (c_i, [c_p, c_p, c_p, c_p]),
    'mmlf_g_pack_weight': (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    'mmlf_g_conv': (c_i, [c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p]),
    'mmlf_g_conv_wgrad': (c_i, [c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    'mmlf_g_colsum': (c_i, [c_p, c_i, c_i, c_i64, c_p, c_p]),
    'mmlf_g_bn_stats': (c_i, [c_p, c_i, c_i, c_i64, c_p, c_p]),
    'mmlf_g_affine': (c_i, [c_p, c_i, c_p, c_p, c_i, c_i64, c_i, c_p, c_i, c_p]),
    'mmlf_g_bn_bwd': (c_i, [c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_i64, c_i, c_i, c_i64, c_p, c_i, c_p, c_p,
                            c_p]),
    'mmlf_g_relu_bwd': (c_i, [c_p, c_i, c_p, c_i, c_i, c_i64, c_p, c_i, c_p]),
    'mmlf_g_maxpool2': (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    'mmlf_g_maxpool2_bwd': (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    'mmlf_g_copy_window': (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i,
                                 c_i, c_p]),
    'mmlf_g_depth_to_space': (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p]),
    'mmlf_g_layout': (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p]),
}

EXPORTS = tuple(_PROTOS)
ABI_VERSION = 5
_lib = None


def lib():
    """The loaded library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(make -C mmlf_b200/csrc).  mmlf_b200 has no CPU / PyTorch fallback.')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.mmlf_abi_version() != ABI_VERSION:
            raise RuntimeError('libmmlf_b200.so ABI version mismatch')
        _lib = l
    return _lib


# kernels launched per C-ABI call (for the launch count reported by bench.py)
_KERNELS_PER_CALL = {'mmlf_g_bn_bwd': 3, 'mmlf_zero': 0, 'mmlf_adam_step_dev': 2, 'mmlf_augment_fill': 0, 'mmlf_augment_patches': 2, 'mmlf_pack_views_split': 2, 'mmlf_conv2x2_wgrad': 2, 'mmlf_conv2x2_wgrad_canonical': 2, 'mmlf_bn_bwd_apply': 2, 'mmlf_head_small_bwd': 2, 'mmlf_shift_taps': 0}
launch_count = 0
_TRACE = os.environ.get('MMLF_TRACE', '0') == '1'      # print every C-ABI call (debugging aid)
_profile = None          # when set to a list, call() appends (name, start_event, end_event)
profile_tag = None       # optional label of the NEXT profiled call (the engine marks narrow / wide convolutions): name#tag


def set_profile(enabled):
    """Record a CUDA-event pair around every C-ABI call on the current stream (used by bench.py for the per-kernel
    time shares).  Returns the previous record list."""
    global _profile
    old = _profile
    _profile = [] if enabled else None
    return old


def call(name, *args):
    """Call an ``int``-returning entry point and raise with the library's error text on failure."""
    global launch_count, profile_tag
    l = lib()
    if _profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(l, name)(*args)
        e1.record()
        _profile.append((name if profile_tag is None else f'{name}#{profile_tag}', e0, e1))
        profile_tag = None
    else:
        rc = getattr(l, name)(*args)
    if _TRACE:
        import sys
        print(f'[mmlf] {name} -> {rc}', file=sys.stderr, flush=True)
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {l.mmlf_last_error().decode()}')
    launch_count += _KERNELS_PER_CALL.get(name, 1)


BF16, FP16 = 0, 1        # MMLF_BF16 / MMLF_FP16 storage codes

_device_ok = False


def require_device():
    """Fail loudly when there is no sm_100 device (no fallback path exists)."""
    global _device_ok
    if not _device_ok:
        call('mmlf_check_device')
        _device_ok = True


class no_gc_during_capture:
    """Context manager around a CUDA-graph capture: the cyclic garbage collector must not run inside it.  A collection that
    happens to free a dead model's CUDAGraph (module <-> engine reference cycles are only freed by the collector) calls
    cudaGraphExecDestroy while this thread is capturing -- "operation not permitted when stream is capturing" -- and
    invalidates the capture (seen on B200: the 70-member ESE capture allocates enough Python objects to trigger one)."""

    def __enter__(self):
        import gc
        self._was = gc.isenabled()
        gc.collect()
        gc.disable()

    def __exit__(self, *exc):
        import gc
        if self._was:
            gc.enable()
        return False
