'''ctypes binding of include/flexdiffuse_b200.h.

This is the only place Python touches the C ABI.  There is no fallback: if the
shared library is missing or a call returns non-zero, a `NativeError` is raised.
Torch is used solely for device memory, streams and dtype bookkeeping.
'''
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional, Sequence

import torch

LIB_NAME = 'libflexdiffuse_b200.so'
LIB_PATH = Path(__file__).resolve().parent / LIB_NAME

FD_ABI_VERSION = 4
FD_DTYPE_F32 = 0
FD_DTYPE_BF16 = 1
FD_BLEND_OK = 0
FD_BLEND_ZERO_DIVISION = 1
FD_BLEND_RANGE = 2

# every symbol include/flexdiffuse_b200.h declares (tests check they all export)
ABI_SYMBOLS = ('fd_version', 'fd_last_error_string', 'fd_arch_check',
               'fd_sm_count', 'fd_cfg_sched_step', 'fd_sim_blend',
               'fd_sim_blend_workspace_bytes',
               'fd_kv_project', 'fd_cross_attn', 'fd_cross_attn_fused',
               'fd_groupnorm_act_workspace_bytes', 'fd_groupnorm_act',
               'fd_add_bias_residual', 'fd_add_layernorm', 'fd_geglu',
               'fd_composite_eps', 'fd_image_tail_u8', 'fd_visual_projection',
               'fd_visual_projection_workspace_bytes', 'fd_visual_projection_range_flag',
               'fd_linear_x3_operand_bytes', 'fd_linear_x3_split', 'fd_linear_x3_split_ln', 'fd_linear_x3',
               'fd_linear_x3_flag',
               'fd_attention_f32', 'fd_ff_geglu', 'fd_concat_channels', 'fd_upsample_nearest2x', 'fd_add_groupnorm_act')


class NativeError(RuntimeError):
    '''A C-ABI call failed (or the library is not built).'''


class SchedCoeffs(C.Structure):
    '''struct fd_sched_coeffs'''
    _fields_ = [('guidance', C.c_float), ('use_cfg', C.c_int),
                ('w', C.c_float * 4), ('a', C.c_float), ('b', C.c_float),
                ('c_noise', C.c_float), ('in_scale', C.c_float)]


class EntityBox(C.Structure):
    '''struct fd_entity_box'''
    _fields_ = [('ox', C.c_int), ('oy', C.c_int), ('sx', C.c_int), ('sy', C.c_int),
                ('blend', C.c_float)]


class TweenParams(C.Structure):
    '''struct fd_tween_params'''
    _fields_ = [('threshold_floor', C.c_double), ('threshold_mult', C.c_double),
                ('clustered', C.c_double), ('max_guidance', C.c_double),
                ('header_max', C.c_double), ('align_mode', C.c_int),
                ('mapping_reuse', C.c_int), ('blend_mode', C.c_int),
                ('reserved', C.c_int)]


BLEND_MODE_LERP = 0
BLEND_MODE_SLERP = 1


_lib: Optional[C.CDLL] = None

# number of kernels of THIS library launched from this process (graph replays are added by
# the caller that owns the graph); bench.py reports it as `gpu_launches`
LAUNCHES = 0


def count_launch(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def lib() -> C.CDLL:
    '''Load (once) and return the native library; raise loudly if absent.'''
    global _lib
    if _lib is not None:
        return _lib
    lib_path = LIB_PATH
    if os.environ.get('FD_LIB_PATH'):  # development aid: A/B a variant build (profiles/build_variants.py)
        lib_path = Path(os.environ['FD_LIB_PATH'])
    if not lib_path.exists():
        raise NativeError(
            f'{lib_path} is not built. Run `python -m flexdiffuse_b200.build` '
            '(needs nvcc). flexdiffuse_b200 has no CPU / PyTorch fallback.')
    l = C.CDLL(str(lib_path))
    l.fd_version.restype = C.c_int
    l.fd_last_error_string.restype = C.c_char_p
    l.fd_arch_check.argtypes = [C.c_int]
    l.fd_arch_check.restype = C.c_int
    l.fd_sm_count.restype = C.c_int
    vp = C.c_void_p
    l.fd_cfg_sched_step.argtypes = [
        vp, vp, C.c_int, vp, vp, vp, vp, vp,
        C.POINTER(SchedCoeffs), C.c_int64, vp, vp, vp, C.c_int, vp
    ]
    l.fd_cfg_sched_step.restype = C.c_int
    l.fd_sim_blend.argtypes = [
        vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int,
        vp, vp, vp, vp, vp, vp, vp, C.c_int64, vp, vp
    ]
    l.fd_sim_blend.restype = C.c_int
    l.fd_sim_blend_workspace_bytes.argtypes = [C.c_int] * 3
    l.fd_sim_blend_workspace_bytes.restype = C.c_int64
    l.fd_kv_project.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    l.fd_kv_project.restype = C.c_int
    l.fd_cross_attn.argtypes = [
        vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, vp, C.c_int, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp
    ]
    l.fd_cross_attn.restype = C.c_int
    l.fd_cross_attn_fused.argtypes = [
        vp, vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, vp, vp, vp, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp
    ]
    l.fd_cross_attn_fused.restype = C.c_int
    l.fd_groupnorm_act_workspace_bytes.argtypes = [C.c_int] * 4
    l.fd_groupnorm_act_workspace_bytes.restype = C.c_int64
    l.fd_add_bias_residual.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int, vp]
    l.fd_add_bias_residual.restype = C.c_int
    l.fd_groupnorm_act.argtypes = [
        vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
        C.c_int, C.c_int64, vp
    ]
    l.fd_groupnorm_act.restype = C.c_int
    l.fd_add_groupnorm_act.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                       C.c_int, vp]
    l.fd_add_groupnorm_act.restype = C.c_int
    l.fd_add_layernorm.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int,
                                   C.c_float, C.c_int64, vp]
    l.fd_add_layernorm.restype = C.c_int
    l.fd_composite_eps.argtypes = [vp, C.c_int, C.POINTER(EntityBox), C.c_int, C.c_int,
                                   C.c_int, C.c_int, vp, vp, vp]
    l.fd_composite_eps.restype = C.c_int
    l.fd_visual_projection_workspace_bytes.argtypes = [C.c_int, C.c_int]
    l.fd_visual_projection_workspace_bytes.restype = C.c_int64
    l.fd_visual_projection.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int64, C.c_int, vp]
    l.fd_visual_projection.restype = C.c_int
    l.fd_visual_projection_range_flag.restype = C.c_int
    l.fd_linear_x3_operand_bytes.argtypes = [C.c_int, C.c_int]
    l.fd_linear_x3_operand_bytes.restype = C.c_int64
    l.fd_linear_x3_split.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int64, vp]
    l.fd_linear_x3_split.restype = C.c_int
    l.fd_linear_x3_split_ln.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_float, vp, C.c_int64, vp]
    l.fd_linear_x3_split_ln.restype = C.c_int
    l.fd_linear_x3.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp]
    l.fd_linear_x3.restype = C.c_int
    l.fd_linear_x3_flag.restype = C.c_int
    l.fd_attention_f32.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_int, vp]
    l.fd_attention_f32.restype = C.c_int
    l.fd_ff_geglu.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int64, vp]
    l.fd_ff_geglu.restype = C.c_int
    l.fd_concat_channels.argtypes = [vp, vp, vp, C.c_int64, C.c_int, C.c_int, vp]
    l.fd_concat_channels.restype = C.c_int
    l.fd_upsample_nearest2x.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    l.fd_upsample_nearest2x.restype = C.c_int
    l.fd_image_tail_u8.argtypes = [vp, C.c_int, C.c_int64, vp, vp]
    l.fd_image_tail_u8.restype = C.c_int
    l.fd_geglu.argtypes = [vp, vp, C.c_int64, C.c_int, vp]
    l.fd_geglu.restype = C.c_int
    if l.fd_version() != FD_ABI_VERSION:
        raise NativeError(f'ABI mismatch: library {l.fd_version()} != '
                          f'binding {FD_ABI_VERSION}; rebuild')
    _lib = l
    return l


def last_error() -> str:
    return lib().fd_last_error_string().decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f'{what} failed (code {rc}): {last_error()}')


def require_device(device: torch.device | int | None = None) -> None:
    '''Raise unless `device` is an sm_100 GPU.  No fallback.'''
    if device is None:
        idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
    elif isinstance(device, int):
        idx = device
    else:
        idx = device.index if device.index is not None else (
            torch.cuda.current_device() if torch.cuda.is_available() else 0)
    check(lib().fd_arch_check(idx), 'fd_arch_check')


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return FD_DTYPE_F32
    if dt == torch.bfloat16:
        return FD_DTYPE_BF16
    raise NativeError(f'unsupported dtype {dt} (need float32 or bfloat16)')


def _need(t: torch.Tensor, name: str, dtype=None) -> None:
    if not t.is_cuda:
        raise NativeError(
            f'{name} must be a CUDA tensor (got {t.device}); flexdiffuse_b200 '
            'has no CPU fallback')
    if not t.is_contiguous():
        raise NativeError(f'{name} must be contiguous')
    if dtype is not None and t.dtype != dtype:
        raise NativeError(f'{name} must be {dtype}, got {t.dtype}')


# --------------------------------------------------------------------------- K4
def cfg_sched_step(eps_uncond: Optional[torch.Tensor],
                   eps_cond: torch.Tensor,
                   x: torch.Tensor,
                   coeffs: SchedCoeffs,
                   x_out: torch.Tensor,
                   hist: Sequence[torch.Tensor] = (),
                   noise: Optional[torch.Tensor] = None,
                   eps_out: Optional[torch.Tensor] = None,
                   scaled_out: Optional[torch.Tensor] = None) -> None:
    '''fd_cfg_sched_step on the current stream of `x.device`.'''
    _need(eps_cond, 'eps_cond')
    _need(x, 'x', torch.float32)
    _need(x_out, 'x_out', torch.float32)
    n = x.numel()
    for name, t in (('eps_uncond', eps_uncond), ('noise', noise),
                    ('eps_out', eps_out), ('scaled_out', scaled_out),
                    *((f'hist[{i}]', h) for i, h in enumerate(hist))):
        if t is not None:
            _need(t, name)
            if t.numel() != n:
                raise NativeError(f'{name} has {t.numel()} elements, x has {n}')
    if eps_cond.numel() != n or x_out.numel() != n:
        raise NativeError('eps_cond / x_out size mismatch with x')
    if eps_uncond is not None and eps_uncond.dtype != eps_cond.dtype:
        raise NativeError('eps_uncond / eps_cond dtype mismatch')
    for h in hist:
        if h.dtype != torch.float32:
            raise NativeError('history tensors must be float32')
    if len(hist) > 3:
        raise NativeError('at most 3 history tensors')
    hp = [ptr(h) for h in hist] + [None] * (3 - len(hist))
    rc = lib().fd_cfg_sched_step(
        ptr(eps_uncond), ptr(eps_cond), dtype_code(eps_cond.dtype), ptr(x),
        hp[0], hp[1], hp[2], ptr(noise), C.byref(coeffs), n, ptr(x_out),
        ptr(eps_out), ptr(scaled_out),
        dtype_code(scaled_out.dtype) if scaled_out is not None else FD_DTYPE_F32,
        stream_ptr(x.device))
    check(rc, 'fd_cfg_sched_step')
    count_launch()


# --------------------------------------------------------------------------- K1
_k1_params = {}      # (device, parameter bytes) -> device copy of the fd_tween_params array
_k1_workspace = {}   # device -> grow-only scratch for the guide's split planes


def sim_blend(text: torch.Tensor,
              guide: torch.Tensor,
              params: Sequence[TweenParams],
              linear_weights: torch.Tensor,
              want_sim: bool = False,
              want_maps: bool = True):
    '''fd_sim_blend.  text [B,T,D] f32, guide [G,A,D] f32 (G in {1,B}),
    linear_weights [P,T] f32 on the same device.
    Returns dict(out [B,P,T,D], map_s [B,P,T], map_idx [B,P,T], weights [B,P,T],
    status [B,P], sim [B,A,T] | None); the three map tensors are None with want_maps=False.
    Host cost per call: the output allocations only -- the parameter array is uploaded once per
    distinct value (pinned, cached) and the workspace is a grow-only per-device buffer.'''
    _need(text, 'text', torch.float32)
    _need(guide, 'guide', torch.float32)
    _need(linear_weights, 'linear_weights', torch.float32)
    B, T, D = text.shape
    G, A, D2 = guide.shape
    P = len(params)
    if D2 != D or G not in (1, B):
        raise NativeError(f'guide shape {tuple(guide.shape)} incompatible with '
                          f'text {tuple(text.shape)}')
    if tuple(linear_weights.shape) != (P, T):
        raise NativeError(f'linear_weights must be [{P},{T}]')
    dev = text.device
    out = torch.empty((B, P, T, D), dtype=torch.float32, device=dev)
    status = torch.empty((B, P), dtype=torch.int32, device=dev)
    map_s = map_idx = weights = None
    if want_maps:
        map_s = torch.empty((B, P, T), dtype=torch.float32, device=dev)
        map_idx = torch.empty((B, P, T), dtype=torch.int32, device=dev)
        weights = torch.empty((B, P, T), dtype=torch.float32, device=dev)
    sim = (torch.empty((B, A, T), dtype=torch.float32, device=dev)
           if want_sim else None)
    arr = (TweenParams * P)(*params)
    raw = bytes(arr)
    key = (dev, raw)
    params_dev = _k1_params.get(key)
    if params_dev is None:
        if len(_k1_params) > 512:
            _k1_params.clear()
        host = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
        params_dev = host.to(dev, non_blocking=True)
        _k1_params[key] = params_dev
    ws_bytes = lib().fd_sim_blend_workspace_bytes(G, A, D)
    ws = _k1_workspace.get(dev)
    if ws is None or ws.numel() < ws_bytes:
        ws = torch.empty(max(ws_bytes, 1 << 22), dtype=torch.uint8, device=dev)
        _k1_workspace[dev] = ws
    rc = lib().fd_sim_blend(ptr(text), ptr(guide), B, G, T, A, D,
                            ptr(params_dev),
                            ptr(linear_weights), P, ptr(out), ptr(map_s),
                            ptr(map_idx), ptr(weights), ptr(status), ptr(sim),
                            ptr(ws), ws.numel(), C.cast(arr, C.c_void_p), stream_ptr(dev))
    count_launch()  # guide prep kernel
    check(rc, 'fd_sim_blend')
    count_launch()
    return dict(out=out, map_s=map_s, map_idx=map_idx, weights=weights,
                status=status, sim=sim)


# --------------------------------------------------------------------------- K2
def kv_project(ctx: torch.Tensor, w: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    '''fd_kv_project: out[M,N] = ctx[M,K] @ w[N,K]^T, all bf16.'''
    _need(ctx, 'ctx', torch.bfloat16)
    _need(w, 'w', torch.bfloat16)
    M, K = ctx.shape
    N, K2 = w.shape
    if K2 != K:
        raise NativeError('ctx / w inner dimension mismatch')
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=ctx.device)
    _need(out, 'out', torch.bfloat16)
    rc = lib().fd_kv_project(ptr(ctx), ptr(w), ptr(out), M, N, K,
                             stream_ptr(ctx.device))
    check(rc, 'fd_kv_project')
    count_launch()
    return out


# --------------------------------------------------------------------------- K3
def cross_attn(q: torch.Tensor, kv: torch.Tensor, k_col_off: int,
               v_col_off: int, ctx_index: torch.Tensor, heads: int,
               t_valid: int, t_pad: int, scale: float,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    '''fd_cross_attn: q [S,Nq,C] bf16, kv = K2 output [n_ctx*t_pad, N] bf16.'''
    _need(q, 'q', torch.bfloat16)
    _need(kv, 'kv', torch.bfloat16)
    _need(ctx_index, 'ctx_index', torch.int32)
    S, Nq, Cc = q.shape
    if Cc % heads:
        raise NativeError('channels not divisible by heads')
    if out is None:
        out = torch.empty_like(q)
    _need(out, 'out', torch.bfloat16)
    rc = lib().fd_cross_attn(ptr(q), ptr(kv), kv.shape[0], kv.shape[1],
                             k_col_off, v_col_off, ptr(ctx_index), S, Nq, heads,
                             Cc // heads, t_valid, t_pad, float(scale),
                             ptr(out), stream_ptr(q.device))
    check(rc, 'fd_cross_attn')
    count_launch()
    return out


# --------------------------------------------------------------------------- K3F
def cross_attn_fused(x: torch.Tensor, wq: torch.Tensor, kv: torch.Tensor,
                     k_col_off: int, v_col_off: int, ctx_index: torch.Tensor,
                     wo: torch.Tensor, bo: torch.Tensor, heads: int, t_valid: int,
                     t_pad: int, scale: float,
                     attn: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None,
                     want_attn: bool = True):
    '''fd_cross_attn_fused: x [S,Nq,C] bf16 -> (out, attn), both [S,Nq,C] bf16, where
    attn = softmax(to_q(x) K^T scale) V over the K2 cache and out = to_out(attn) + bias.
    `want_attn=False` with C == 320 keeps the attention output on chip (attn is None); wider
    layers always need the buffer (their head groups exchange it through L2).'''
    _need(x, 'x', torch.bfloat16)
    _need(wq, 'wq', torch.bfloat16)
    _need(wo, 'wo', torch.bfloat16)
    _need(bo, 'bo', torch.bfloat16)
    _need(kv, 'kv', torch.bfloat16)
    _need(ctx_index, 'ctx_index', torch.int32)
    S, Nq, Cc = x.shape
    if Cc % heads or tuple(wq.shape) != (Cc, Cc) or tuple(wo.shape) != (Cc, Cc) \
            or bo.numel() != Cc:
        raise NativeError('cross_attn_fused: weight shapes do not match x')
    if attn is None and (want_attn or Cc != 320):
        attn = torch.empty_like(x)
    if out is None:
        out = torch.empty_like(x)
    if attn is not None:
        _need(attn, 'attn', torch.bfloat16)
    _need(out, 'out', torch.bfloat16)
    rc = lib().fd_cross_attn_fused(ptr(x), ptr(wq), ptr(kv), kv.shape[0], kv.shape[1],
                                   k_col_off, v_col_off, ptr(ctx_index), ptr(wo),
                                   ptr(bo), S, Nq, heads, Cc // heads, t_valid, t_pad,
                                   float(scale), ptr(attn), ptr(out),
                                   stream_ptr(x.device))
    check(rc, 'fd_cross_attn_fused')
    count_launch()
    return out, attn


def k3f_status(reset: bool = True):
    '''Development aid: watchdog record of the fused kernel (all zeros = no wait timed out).'''
    arr = (C.c_int * 4)()
    check(lib().fd_debug_k3f_status(arr, int(reset)), 'fd_debug_k3f_status')
    return list(arr)


# --------------------------------------------------------------------------- K5 / K6
_gn_workspaces = {}
_gn_retired = []


def groupnorm_act(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  groups: int, eps: float, silu: bool,
                  bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    '''fd_groupnorm_act on a channels-last bf16 [N,C,H,W] tensor -> same layout.'''
    if x.dtype != torch.bfloat16 or not x.is_cuda:
        raise NativeError('groupnorm_act needs a CUDA bfloat16 tensor '
                          f'(got {x.dtype} on {x.device}); no fallback')
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    N, Cc, H, W = x.shape
    bias_stride = 0
    if bias is not None:
        if (not bias.is_cuda or bias.dtype != torch.bfloat16
                or tuple(bias.shape) != (N, Cc) or bias.stride(1) != 1):
            raise NativeError(f'bias must be CUDA bfloat16 [{N},{Cc}] with unit '
                              'column stride')
        bias_stride = bias.stride(0) if N > 1 else Cc
    y = torch.empty_like(x)  # preserves channels_last
    nbytes = lib().fd_groupnorm_act_workspace_bytes(N, H * W, Cc, groups)
    ws = _gn_workspaces.get(x.device)
    if ws is None or ws.numel() < nbytes:
        # one grow-only scratch per device; launches on a stream are ordered, so reuse is safe
        if ws is not None:
            _gn_retired.append(ws)  # a captured CUDA graph may still point at it
        ws = torch.zeros(max(2 * nbytes, 1 << 24), dtype=torch.uint8, device=x.device)
        _gn_workspaces[x.device] = ws
    rc = lib().fd_groupnorm_act(ptr(x), ptr(bias), ptr(gamma), ptr(beta), ptr(ws),
                                ptr(y), N, H * W, Cc, groups, float(eps),
                                int(silu), bias_stride, stream_ptr(x.device))
    check(rc, 'fd_groupnorm_act')
    count_launch(1 if x.numel() * 2 <= (8 << 20) else 3)  # cluster kernel or stats/finalize/apply
    return y


def _gn_workspace(device, nbytes: int) -> torch.Tensor:
    ws = _gn_workspaces.get(device)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _gn_retired.append(ws)  # a captured CUDA graph may still point at it
        ws = torch.zeros(max(2 * nbytes, 1 << 24), dtype=torch.uint8, device=device)
        _gn_workspaces[device] = ws
    return ws


def add_groupnorm_act(x: torch.Tensor, h: torch.Tensor, rbias: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                      groups: int, eps: float, silu: bool):
    '''fd_add_groupnorm_act on channels-last bf16 [N,C,H,W] tensors: returns (x + h + rbias[c], act(GroupNorm(that))).'''
    for name, t in (('x', x), ('h', h)):
        if t.dtype != torch.bfloat16 or not t.is_cuda or t.dim() != 4:
            raise NativeError(f'add_groupnorm_act: {name} must be a CUDA bfloat16 [N,C,H,W] tensor; no fallback')
    if x.shape != h.shape:
        raise NativeError('add_groupnorm_act: x / h shape mismatch')
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    if not h.is_contiguous(memory_format=torch.channels_last):
        h = h.contiguous(memory_format=torch.channels_last)
    _need(rbias, 'rbias', torch.bfloat16)
    N, Cc, H, W = x.shape
    total, y = torch.empty_like(x), torch.empty_like(x)
    ws = _gn_workspace(x.device, lib().fd_groupnorm_act_workspace_bytes(N, H * W, Cc, groups))
    check(lib().fd_add_groupnorm_act(ptr(x), ptr(h), ptr(rbias), ptr(total), ptr(gamma), ptr(beta), ptr(ws), ptr(y), N, H * W,
                                     Cc, groups, float(eps), int(silu), stream_ptr(x.device)), 'fd_add_groupnorm_act')
    count_launch(1 if x.numel() * 2 <= (8 << 20) else 4)
    return total, y


def geglu(x: torch.Tensor) -> torch.Tensor:
    '''fd_geglu: [..., 2F] bf16 -> [..., F].'''
    _need(x, 'x', torch.bfloat16)
    F2 = x.shape[-1]
    out = torch.empty(x.shape[:-1] + (F2 // 2,), dtype=x.dtype, device=x.device)
    rc = lib().fd_geglu(ptr(x), ptr(out), x.numel() // F2, F2 // 2,
                        stream_ptr(x.device))
    check(rc, 'fd_geglu')
    count_launch()
    return out


def add_bias_residual(x: torch.Tensor, h: Optional[torch.Tensor],
                      bias: torch.Tensor, inplace: bool = False) -> torch.Tensor:
    '''fd_add_bias_residual: x + h + bias[c] on channels-last bf16 [N,C,H,W] tensors; h None: x + bias[c]
    (`inplace`: written over x when x is already channels-last).'''
    for name, t in (('x', x), ('h', h)):
        if t is not None and (t.dtype != torch.bfloat16 or not t.is_cuda):
            raise NativeError(f'{name} must be CUDA bfloat16; no fallback')
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    if h is not None and not h.is_contiguous(memory_format=torch.channels_last):
        h = h.contiguous(memory_format=torch.channels_last)
    _need(bias, 'bias', torch.bfloat16)
    y = x if inplace else torch.empty_like(x)
    rc = lib().fd_add_bias_residual(ptr(x), ptr(h), ptr(bias), ptr(y), x.numel(),
                                    x.shape[1], stream_ptr(x.device))
    check(rc, 'fd_add_bias_residual')
    count_launch()
    return y


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    '''fd_concat_channels: torch.cat([a, b], dim=1) for channels-last bf16 [N,C,H,W] tensors (C % 8 == 0).'''
    for name, t in (('a', a), ('b', b)):
        if t.dtype != torch.bfloat16 or not t.is_cuda or t.dim() != 4:
            raise NativeError(f'{name} must be a CUDA bfloat16 [N,C,H,W] tensor; no fallback')
    if a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        raise NativeError(f'concat_channels: shapes {tuple(a.shape)} and {tuple(b.shape)} do not match')
    if not a.is_contiguous(memory_format=torch.channels_last):
        a = a.contiguous(memory_format=torch.channels_last)
    if not b.is_contiguous(memory_format=torch.channels_last):
        b = b.contiguous(memory_format=torch.channels_last)
    N, Ca, H, W = a.shape
    Cb = b.shape[1]
    y = torch.empty((N, Ca + Cb, H, W), dtype=torch.bfloat16, device=a.device, memory_format=torch.channels_last)
    check(lib().fd_concat_channels(ptr(a), ptr(b), ptr(y), N * H * W, Ca, Cb, stream_ptr(a.device)), 'fd_concat_channels')
    count_launch()
    return y


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    '''fd_upsample_nearest2x: F.interpolate(x, scale_factor=2.0, mode='nearest') for a channels-last bf16 [N,C,H,W] tensor.'''
    if x.dtype != torch.bfloat16 or not x.is_cuda or x.dim() != 4:
        raise NativeError('upsample_nearest2x needs a CUDA bfloat16 [N,C,H,W] tensor; no fallback')
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    N, Cc, H, W = x.shape
    y = torch.empty((N, Cc, 2 * H, 2 * W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    check(lib().fd_upsample_nearest2x(ptr(x), ptr(y), N, H, W, Cc, stream_ptr(x.device)), 'fd_upsample_nearest2x')
    count_launch()
    return y


def _column_block(t: torch.Tensor, rows: int, cols: int, name: str) -> int:
    '''Row stride (elements) of `t`, a [rows, cols] bf16 column block of a wider row-major matrix.'''
    if (t.dtype != torch.bfloat16 or not t.is_cuda or t.dim() != 2 or tuple(t.shape) != (rows, cols) or t.stride(1) != 1
            or t.stride(0) < cols or t.data_ptr() % 16):
        raise NativeError(f'{name} must be a CUDA bfloat16 [{rows}, {cols}] view with unit column stride, 16-byte aligned')
    return t.stride(0)


def add_layernorm(x: torch.Tensor, y: Optional[torch.Tensor], gamma: torch.Tensor,
                  beta: torch.Tensor, eps: float, sum_out: Optional[torch.Tensor] = None):
    '''fd_add_layernorm.  Returns (x + y, LayerNorm(x + y)); with y None: (x, LayerNorm(x)).  `sum_out`: a [M, C]
    column block of a wider matrix to receive x + y instead of a fresh tensor.'''
    _need(x, 'x', torch.bfloat16)
    Cc = x.shape[-1]
    norm = torch.empty_like(x)
    total, stride = None, 0
    if y is not None:
        _need(y, 'y', torch.bfloat16)
        if y.shape != x.shape:
            raise NativeError('x / y shape mismatch')
        if sum_out is not None:
            stride = _column_block(sum_out, x.numel() // Cc, Cc, 'sum_out')
            total = sum_out
        else:
            total = torch.empty_like(x)
    elif sum_out is not None:
        raise NativeError('add_layernorm: sum_out without y')
    rc = lib().fd_add_layernorm(ptr(x), ptr(y), ptr(gamma), ptr(beta), ptr(total),
                                ptr(norm), x.numel() // Cc, Cc, float(eps), stride,
                                stream_ptr(x.device))
    check(rc, 'fd_add_layernorm')
    count_launch()
    return (total if y is not None else x), norm


def composite_eps(eps_all: torch.Tensor, boxes: Sequence[EntityBox]):
    '''fd_composite_eps: eps_all [2+E, C, H, W] (uncond, background, entities; NCHW contiguous)
    -> (uncond fp32 [1,C,H,W], composite conditional fp32 [1,C,H,W]).'''
    _need(eps_all, 'eps_all')
    n, Cc, H, W = eps_all.shape
    if n != 2 + len(boxes):
        raise NativeError(f'eps_all has {n} samples, expected {2 + len(boxes)}')
    u = torch.empty((1, Cc, H, W), dtype=torch.float32, device=eps_all.device)
    c = torch.empty_like(u)
    arr = (EntityBox * max(len(boxes), 1))(*boxes)
    rc = lib().fd_composite_eps(ptr(eps_all), dtype_code(eps_all.dtype), arr, len(boxes),
                                Cc, H, W, ptr(u), ptr(c), stream_ptr(eps_all.device))
    check(rc, 'fd_composite_eps')
    count_launch()
    return u, c


# --------------------------------------------------------------------------- K10
def image_tail_u8(image: torch.Tensor) -> torch.Tensor:
    '''fd_image_tail_u8: decoder output [B,3,H,W] (f32 / bf16) -> uint8 [B,H,W,3] =
    round(clamp(x / 2 + 0.5, 0, 1) * 255), the bytes PIL gets in flex.py:119-124.'''
    if not image.is_cuda:
        raise NativeError('image_tail_u8 needs a CUDA tensor; no fallback')
    if not image.is_contiguous(memory_format=torch.channels_last):
        image = image.contiguous(memory_format=torch.channels_last)
    B, Cc, H, W = image.shape
    out = torch.empty((B, H, W, Cc), dtype=torch.uint8, device=image.device)
    rc = lib().fd_image_tail_u8(ptr(image), dtype_code(image.dtype), image.numel(), ptr(out),
                                stream_ptr(image.device))
    check(rc, 'fd_image_tail_u8')
    count_launch()
    return out


# --------------------------------------------------------------------------- K1P
_proj_planes = {}


def visual_projection(hidden: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    '''fd_visual_projection: hidden [..., K] fp32 @ weight[N, K]^T -> [..., N] fp32 with fp32
    accuracy on tcgen05.  The split weight planes are cached per (weight storage, version);
    the cache holds a reference to the weight.'''
    _need(weight, 'weight', torch.float32)
    if hidden.dtype != torch.float32 or not hidden.is_cuda:
        raise NativeError('visual_projection needs CUDA float32 hidden states')
    h2 = hidden.reshape(-1, hidden.shape[-1]).contiguous()
    M, K = h2.shape
    N = weight.shape[0]
    key = (weight.data_ptr(), weight._version, N, K)
    ent = _proj_planes.get(weight.device)
    changed = ent is None or ent[0] != key
    if changed:
        ws = torch.empty(lib().fd_visual_projection_workspace_bytes(N, K), dtype=torch.uint8,
                         device=weight.device)
        # the entry keeps the weight tensor alive: its address cannot be handed to another tensor while
        # the planes made from it are cached (a freed weight's address coming back with version 0 and
        # the same shape would otherwise hit the cache with stale planes)
        _proj_planes[weight.device] = (key, ws, weight)
    ws = _proj_planes[weight.device][1]
    out = torch.empty((M, N), dtype=torch.float32, device=hidden.device)
    rc = lib().fd_visual_projection(ptr(h2), ptr(weight), ptr(out), M, N, K, ptr(ws), ws.numel(),
                                    int(changed), stream_ptr(hidden.device))
    check(rc, 'fd_visual_projection')
    count_launch(2 if changed else 1)
    return out.reshape(*hidden.shape[:-1], N)


# --------------------------------------------------------------------------- K11
LINEAR_ACT_NONE, LINEAR_ACT_QUICK_GELU, LINEAR_ACT_GELU = 0, 1, 2
_x3_weights = {}     # (data_ptr, version, N, K) -> (operand buffer, the weight tensor it was made from)


def x3_split(x2d: torch.Tensor) -> torch.Tensor:
    '''fd_linear_x3_split: fp32 [rows, K] -> operand buffer (two fp16 planes + inverse row scales).'''
    _need(x2d, 'x', torch.float32)
    rows, K = x2d.shape
    buf = torch.empty(lib().fd_linear_x3_operand_bytes(rows, K), dtype=torch.uint8, device=x2d.device)
    check(lib().fd_linear_x3_split(ptr(x2d), rows, K, ptr(buf), buf.numel(), stream_ptr(x2d.device)),
          'fd_linear_x3_split')
    count_launch()
    return buf


def x3_split_ln(x2d: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> torch.Tensor:
    '''fd_linear_x3_split_ln: operand buffer of LayerNorm(x2d) * gamma + beta (K % 64 == 0, K <= 8192).'''
    _need(x2d, 'x', torch.float32)
    _need(gamma, 'gamma', torch.float32)
    _need(beta, 'beta', torch.float32)
    rows, K = x2d.shape
    buf = torch.empty(lib().fd_linear_x3_operand_bytes(rows, K), dtype=torch.uint8, device=x2d.device)
    check(lib().fd_linear_x3_split_ln(ptr(x2d), rows, K, ptr(gamma), ptr(beta), float(eps), ptr(buf), buf.numel(),
                                      stream_ptr(x2d.device)), 'fd_linear_x3_split_ln')
    count_launch()
    return buf


def x3_weight_operand(weight: torch.Tensor) -> torch.Tensor:
    '''Split planes of a weight matrix, cached per (storage, version); the cache keeps the weight alive so
    its address cannot be recycled under a stale entry.'''
    key = (weight.data_ptr(), weight._version, tuple(weight.shape))
    ent = _x3_weights.get(key)
    if ent is None:
        if len(_x3_weights) > 1024:
            _x3_weights.clear()
        ent = (x3_split(weight.detach()), weight)
        _x3_weights[key] = ent
    return ent[0]


def linear_x3(x: Optional[torch.Tensor], weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
              act: int = LINEAR_ACT_NONE, operand: Optional[torch.Tensor] = None,
              split_k: Optional[int] = None, residual: Optional[torch.Tensor] = None,
              rows: Optional[int] = None) -> torch.Tensor:
    '''fd_linear_x3: residual + act(x[..., K] @ weight[N, K]^T + bias) in fp32 accuracy on tcgen05.  `operand` is
    the result of `x3_split` / `x3_split_ln` on the same (flattened) input, for callers that feed one input to
    several Linears or never materialise it (then pass x=None and `rows`).  Returns [..., N] ([rows, N] without x).'''
    _need(weight, 'weight', torch.float32)
    N, K = weight.shape
    if x is not None:
        if x.dtype != torch.float32 or not x.is_cuda:
            raise NativeError('linear_x3 needs CUDA float32 input')
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        if x2.shape[1] != K:
            raise NativeError(f'linear_x3: weight is {tuple(weight.shape)}, input has K={x2.shape[1]}')
        if operand is None:
            operand = x3_split(x2)
    else:
        if operand is None or rows is None:
            raise NativeError('linear_x3: without x, pass the split operand and its row count')
        M = rows
    dev = weight.device
    if bias is not None:
        _need(bias, 'bias', torch.float32)
    wop = x3_weight_operand(weight)
    if split_k is None:
        # measured on the tower shapes (profiles/k11_bench.py): a second K slice pays while the grid is under
        # ~100 CTAs; long K wants 4 (M > 128) or 8 slices
        if K >= 3072:
            split_k = 4 if M > 128 else 8
        elif K >= 768 and N <= 3072:
            split_k = 2
        else:
            split_k = 1
    out = torch.empty((M, N), dtype=torch.float32, device=dev)
    if residual is not None:
        _need(residual, 'residual', torch.float32)
        if residual.numel() != M * N:
            raise NativeError(f'linear_x3: residual has {residual.numel()} elements, need {M * N}')
    check(lib().fd_linear_x3(ptr(operand), M, ptr(wop), N, K, ptr(bias), act, ptr(residual), ptr(out), split_k,
                             stream_ptr(dev)), 'fd_linear_x3')
    count_launch()
    return out.reshape(*x.shape[:-1], N) if x is not None else out


# --------------------------------------------------------------------------- K12
def attention_f32_supported(T: int, d: int) -> bool:
    '''Shapes K12 serves: K^T, V, the query rows and the probability strips of one head fit in shared memory.'''
    ni = 3 if T <= 96 else 9
    floats = d * (32 * ni + 1) + T * d + 8 * d * 4 + 8 * 32 * ni * 4
    return T <= 288 and d % 4 == 0 and d <= 128 and floats * 4 <= 227 * 1024


def attention_f32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float,
                  causal: bool = False) -> torch.Tensor:
    '''fd_attention_f32.  q / k / v: [B, T, C] fp32 views with unit stride along C and one common token
    stride (e.g. the three column blocks of a fused q|k|v projection); returns [B, T, C] contiguous.'''
    B, T, Cc = q.shape
    for t, name in ((q, 'q'), (k, 'k'), (v, 'v')):
        if not t.is_cuda or t.dtype != torch.float32:
            raise NativeError(f'attention_f32: {name} must be CUDA float32')
        if t.shape != q.shape or t.stride(2) != 1 or t.stride(1) != q.stride(1) or t.stride(0) != T * q.stride(1):
            raise NativeError(f'attention_f32: {name} must share q\'s [B, T, C] layout (unit channel stride)')
    out = torch.empty((B, T, Cc), dtype=torch.float32, device=q.device)
    check(lib().fd_attention_f32(ptr(q), ptr(k), ptr(v), q.stride(1), ptr(out), B, T, heads, Cc // heads,
                                 float(scale), int(causal), stream_ptr(q.device)), 'fd_attention_f32')
    count_launch()
    return out


# --------------------------------------------------------------------------- K13
def ff_geglu(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    '''fd_ff_geglu: x [..., K] bf16, weight [2F, K] bf16 (value rows, gate rows), bias [2F] bf16 ->
    (x W_v^T + b_v) * gelu(x W_g^T + b_g) as [..., F] bf16; the [.., 2F] projection is never materialised.
    `out`: a [M, F] column block of a wider matrix to write into (returned as is).'''
    _need(weight, 'weight', torch.bfloat16)
    _need(bias, 'bias', torch.bfloat16)
    if not x.is_cuda or x.dtype != torch.bfloat16:
        raise NativeError('ff_geglu needs CUDA bfloat16 input')
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, K = x2.shape
    F2 = weight.shape[0]
    if weight.shape[1] != K or F2 % 2 or bias.numel() != F2:
        raise NativeError(f'ff_geglu: weight {tuple(weight.shape)} / bias {tuple(bias.shape)} do not fit K={K}')
    stride = 0
    if out is not None:
        stride = _column_block(out, M, F2 // 2, 'out')
        if stride % 8:
            raise NativeError('ff_geglu: output row stride must be a multiple of 8 elements')
    else:
        out = torch.empty((M, F2 // 2), dtype=torch.bfloat16, device=x.device)
    check(lib().fd_ff_geglu(ptr(x2), ptr(weight), ptr(bias), ptr(out), M, F2 // 2, K, stride, stream_ptr(x.device)),
          'fd_ff_geglu')
    count_launch()
    return out if stride else out.reshape(*x.shape[:-1], F2 // 2)


def ff_geglu_supported(K: int, F: int) -> bool:
    return K % 64 == 0 and F % 128 == 0
