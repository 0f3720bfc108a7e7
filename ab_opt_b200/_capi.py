"""ctypes binding of libabopt_b200.so (the C ABI declared in include/abopt_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded every entry
point raises.  torch is used only to obtain device pointers and the current CUDA stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ABOPT_LIB') or os.path.join(_HERE, '_lib', 'libabopt_b200.so')      # ABOPT_LIB: a variant build (A/B measurements)

OK = 0
SCOPE_FULL, SCOPE_ENCODER, SCOPE_EPSNET = 0, 1, 2
SAMPLE_STRUCTURE, SAMPLE_SEQUENCE, KEEP_TRAJECTORY, GRAD_SEMANTICS = 1, 2, 4, 8

# every symbol include/abopt_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    'abopt_version', 'abopt_last_error', 'abopt_kernel_launch_count', 'abopt_model_create',
    'abopt_model_destroy', 'abopt_model_set_tensor', 'abopt_model_finalize', 'abopt_ga_block_forward',
    'abopt_ga_encoder_forward', 'abopt_ga_block_taps', 'abopt_eps_net_forward', 'abopt_rot_denoise',
    'abopt_pos_pred_noise_from_start', 'abopt_pos_denoise', 'abopt_seq_denoise', 'abopt_sample_device',
    'abopt_sample_host', 'abopt_workspace_bytes', 'abopt_sample_init', 'abopt_reverse_step', 'abopt_profile_enable', 'abopt_profile_collect', 'abopt_debug_gemm3x', 'abopt_debug_clocks', 'abopt_debug_copy',
    'abopt_loss_forward', 'abopt_model_set_batch_offset', 'abopt_pair_embed_create', 'abopt_pair_embed_destroy', 'abopt_pair_embed_set_tensor',
    'abopt_pair_embed_finalize', 'abopt_pair_embed_forward', 'abopt_res_embed_create', 'abopt_res_embed_destroy',
    'abopt_res_embed_set_tensor', 'abopt_res_embed_finalize', 'abopt_res_embed_forward',
    'abopt_reconstruct_backbone_partially', 'abopt_pairwise_rmsd', 'abopt_rank_commoness', 'abopt_design_device', 'abopt_design_host', 'abopt_loss_backward', 'abopt_model_get_grad', 'abopt_ga_block_backward',
)


class Config(C.Structure):
    _fields_ = [('num_layers', C.c_int32), ('num_steps', C.c_int32), ('has_prmsd', C.c_int32),
                ('prmsd_bins', C.c_int32), ('prmsd_min', C.c_float), ('prmsd_max', C.c_float),
                ('obj_pred_x0', C.c_int32), ('scope', C.c_int32)]


class StepNoise(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ('u', 'expo_ang', 'unif_ang', 'gauss_ang', 'z_pos', 'expo_seq')]


class InitNoise(C.Structure):
    _fields_ = [('g4', C.c_void_p), ('gp', C.c_void_p), ('s_rand', C.c_void_p), ('add', C.POINTER(StepNoise))]


class AboptError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AboptError(f'{LIB_PATH} not found: build it with `python -m ab_opt_b200.build` '
                             '(there is no CPU fallback)')
        L = C.CDLL(LIB_PATH)
        L.abopt_last_error.restype = C.c_char_p
        L.abopt_kernel_launch_count.restype = C.c_uint64
        L.abopt_workspace_bytes.restype = C.c_size_t
        L.abopt_workspace_bytes.argtypes = [C.c_void_p]
        L.abopt_model_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]
        L.abopt_model_destroy.argtypes = [C.c_void_p]
        L.abopt_model_destroy.restype = None
        L.abopt_model_set_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.abopt_model_finalize.argtypes = [C.c_void_p]
        L.abopt_model_set_batch_offset.argtypes = [C.c_void_p, C.c_int64]
        vp, ci = C.c_void_p, C.c_int
        L.abopt_ga_block_forward.argtypes = [vp, ci, ci, ci] + [vp] * 7
        L.abopt_ga_encoder_forward.argtypes = [vp, ci, ci] + [vp] * 7
        L.abopt_ga_block_taps.argtypes = [vp, ci, ci, ci] + [vp] * 8
        L.abopt_eps_net_forward.argtypes = [vp, ci, ci] + [vp] * 14
        L.abopt_rot_denoise.argtypes = [vp, ci, ci] + [vp] * 10
        L.abopt_pos_pred_noise_from_start.argtypes = [vp, ci, ci] + [vp] * 6
        L.abopt_pos_denoise.argtypes = [vp, ci, ci] + [vp] * 7
        L.abopt_seq_denoise.argtypes = [vp, ci, ci] + [vp] * 8
        L.abopt_sample_device.argtypes = [vp, ci, ci] + [vp] * 7 + [C.c_uint32, ci, C.c_uint64,
                                                                   C.POINTER(InitNoise), C.POINTER(StepNoise)] + [vp] * 6
        L.abopt_sample_init.argtypes = [vp, ci, ci] + [vp] * 4 + [C.c_uint32, ci, C.c_uint64, C.POINTER(InitNoise)] + [vp] * 4
        L.abopt_reverse_step.argtypes = [vp, ci, ci, ci, ci] + [vp] * 7 + [C.c_uint32, C.c_uint64, C.POINTER(StepNoise)] + [vp] * 6
        L.abopt_debug_gemm3x.argtypes = [ci, ci, ci, ci] + [vp] * 5
        L.abopt_debug_copy.argtypes = [vp, ci, vp, C.c_size_t, C.POINTER(C.c_size_t), vp]
        L.abopt_loss_forward.argtypes = [vp, ci, ci] + [vp] * 7 + [C.c_uint32, vp, C.c_uint64, C.POINTER(StepNoise), vp, vp]
        L.abopt_loss_backward.argtypes = [vp, ci, ci] + [vp] * 7 + [C.c_uint32, vp, C.c_uint64, C.POINTER(StepNoise), vp, vp, vp, vp, vp]
        L.abopt_model_get_grad.argtypes = [vp, C.c_char_p, vp, C.c_size_t, vp]
        L.abopt_ga_block_backward.argtypes = [vp, ci, ci, ci] + [vp] * 9
        L.abopt_sample_host.argtypes = [vp, ci, ci] + [vp] * 7 + [C.c_uint32, ci, C.c_uint64] + [vp] * 5
        L.abopt_pair_embed_create.argtypes = [ci, ci, C.POINTER(C.c_void_p)]
        L.abopt_pair_embed_destroy.argtypes = [vp]
        L.abopt_pair_embed_destroy.restype = None
        L.abopt_pair_embed_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_size_t, ci]
        L.abopt_pair_embed_finalize.argtypes = [vp]
        L.abopt_pair_embed_forward.argtypes = [vp, ci, ci, ci] + [vp] * 9
        L.abopt_res_embed_create.argtypes = [ci, ci, C.POINTER(C.c_void_p)]
        L.abopt_res_embed_destroy.argtypes = [vp]
        L.abopt_res_embed_destroy.restype = None
        L.abopt_res_embed_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_size_t, ci]
        L.abopt_res_embed_finalize.argtypes = [vp]
        L.abopt_res_embed_forward.argtypes = [vp, ci, ci, ci] + [vp] * 10
        L.abopt_reconstruct_backbone_partially.argtypes = [ci, ci, ci] + [vp] * 13
        L.abopt_pairwise_rmsd.argtypes = [ci, ci] + [vp] * 5
        L.abopt_rank_commoness.argtypes = [ci, ci, vp, ci, vp, vp, vp]
        L.abopt_design_device.argtypes = [vp, vp, vp, ci, ci, ci] + [vp] * 8 + [C.c_uint32, ci, C.c_uint64] + [vp] * 6
        L.abopt_design_host.argtypes = [vp, vp, vp, ci, ci, ci] + [vp] * 8 + [C.c_uint32, ci, C.c_uint64] + [vp] * 5
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise AboptError(f'libabopt_b200 error {rc}: {lib().abopt_last_error().decode()}')


KERNEL_KINDS = ('mixer', 'proj', 'logits', 'pair', 'aggr', 'tail', 'heads', 'step', 'other', 'ctx', 'pair_part')


def profile_enable(on):
    check(lib().abopt_profile_enable(int(bool(on))))


def profile_collect():
    """-> {kind: (total_ms, launches)} since profile_enable(True)."""
    n = len(KERNEL_KINDS)
    ms = (C.c_double * n)()
    cnt = (C.c_uint64 * n)()
    check(lib().abopt_profile_collect(ms, cnt, n))
    return {k: (ms[i], int(cnt[i])) for i, k in enumerate(KERNEL_KINDS)}


def launch_count():
    return int(lib().abopt_kernel_launch_count())


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_DTYPE = {torch.float32: 0, torch.bool: 1, torch.uint8: 1, torch.int64: 2}


class NativeModel:
    """Owns one abopt_model handle: weights packed on the device + scratch workspace."""

    def __init__(self, cfg: Config, device, tensors):
        L = lib()
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise AboptError('ab_opt_b200 runs on CUDA devices only (no CPU fallback); got device ' + str(dev))
        self.device = dev
        self.index = dev.index if dev.index is not None else torch.cuda.current_device()
        self.cfg = cfg
        h = C.c_void_p()
        check(L.abopt_model_create(C.byref(cfg), self.index, C.byref(h)))
        self.handle = h
        try:
            for key, t in tensors.items():
                t = t.detach()
                if t.dtype not in _DTYPE:
                    t = t.float()
                t = t.contiguous()
                on_dev = 1 if t.is_cuda else 0
                check(L.abopt_model_set_tensor(h, key.encode(), ptr(t) if t.numel() else None, t.numel(),
                                               _DTYPE[t.dtype], on_dev))
            check(L.abopt_model_finalize(h))
        except Exception:
            L.abopt_model_destroy(h)
            self.handle = None
            raise

    def __del__(self):
        if getattr(self, 'handle', None) is not None and _lib is not None:
            _lib.abopt_model_destroy(self.handle)
            self.handle = None

    def workspace_bytes(self):
        return int(lib().abopt_workspace_bytes(self.handle))


def cuda_f32(t, name):
    if not t.is_cuda:
        raise AboptError(f'{name} must be a CUDA tensor (no CPU fallback)')
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def cuda_mask(t, name):
    if not t.is_cuda:
        raise AboptError(f'{name} must be a CUDA tensor (no CPU fallback)')
    return t.to(torch.bool).contiguous()


def cuda_i64(t, name):
    if not t.is_cuda:
        raise AboptError(f'{name} must be a CUDA tensor (no CPU fallback)')
    return t.to(torch.int64).contiguous()
