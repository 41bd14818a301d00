"""Stage-level wrappers over the C ABI (include/mi_b200.h) for torch CUDA tensors.

PyTorch is used here for device memory, streams and dtype bookkeeping only; every FLOP of the hot
path runs in libmi_b200.so.  All functions raise ``MIError`` on failure — there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

CRITIC = {"dot": 0, "bilinear": 1}
ESTIMATOR = {"dv": 0, "infonce": 1, "infonce_ref": 1, "infonce_row": 2, "infonce_sym": 3}
PRECISION = {"fast": 0, "strict": 1}


class MIError(RuntimeError):
    pass


def _check(status: int, what: str) -> None:
    if status != 0:
        lib = _lib.load()
        msg = lib.mi_status_string(status).decode()
        if status == -3:
            msg += ": " + lib.mi_last_cuda_error().decode()
        raise MIError(f"{what} failed: {msg}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise MIError("mi_b200 has no CPU path: tensors must live on a CUDA (sm_100) device")


_workspaces = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Persistent per-device scratch, grown on demand (stream-ordered reuse inside the library)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        _workspaces[key] = None
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=f"cuda:{key}")
        _workspaces[key] = buf
    return buf


def as_bf16(x: torch.Tensor) -> torch.Tensor:
    """bf16, contiguous copy/cast through the library's cast kernel for fp32 inputs."""
    _need_cuda(x)
    if x.dtype == torch.bfloat16:
        return x.contiguous()
    x = x.detach().contiguous().float()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if x.numel():
        _check(_lib.load().mi_cast_f32_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), "mi_cast_f32_to_bf16")
    return out


def gemm(A: torch.Tensor, B: torch.Tensor, alpha: float = 1.0, gamma: float = 0.0,
         sub: Optional[torch.Tensor] = None, out_dtype=torch.float32) -> torch.Tensor:
    """C = alpha * (A @ B.T - gamma * sub); A [M,K], B [N,K] bf16."""
    _need_cuda(A, B, sub)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.shape[1] == B.shape[1]
    A = A if A.stride(1) == 1 and A.stride(0) % 8 == 0 else A.contiguous()     # row-strided views are fine
    B = B if B.stride(1) == 1 and B.stride(0) % 8 == 0 else B.contiguous()
    M, K = A.shape
    N = B.shape[0]
    if A.stride(0) % 8 or B.stride(0) % 8:
        raise MIError("gemm operands need a row pitch that is a multiple of 8 elements (TMA: 16-byte strides)")
    out = torch.empty((M, N), dtype=out_dtype, device=A.device)
    of = _ptr(out) if out_dtype == torch.float32 else None
    ob = _ptr(out) if out_dtype == torch.bfloat16 else None
    if sub is not None:
        sub = sub.contiguous()
    _check(_lib.load().mi_gemm_bf16(_ptr(A), A.stride(0), _ptr(B), B.stride(0), M, N, K, alpha, gamma, _ptr(sub),
                                    0 if sub is None else sub.shape[1], of, ob, N, None, 0, _stream()), "mi_gemm_bf16")
    return out


def transpose(x: torch.Tensor, ld_out: Optional[int] = None) -> torch.Tensor:
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 2
    x = x.contiguous()
    R, Cc = x.shape
    ld = ld_out or ((R + 7) // 8) * 8          # pitch padded so the result can feed the TMA-based GEMM
    out = torch.zeros((Cc, ld), dtype=torch.bfloat16, device=x.device)
    _check(_lib.load().mi_transpose_bf16(_ptr(x), Cc, _ptr(out), ld, R, Cc, _stream()), "mi_transpose_bf16")
    return out[:, :R]


def score_stats(Q: torch.Tensor, K: torch.Tensor, sid_q: torch.Tensor, sid_k: torch.Tensor,
                q_offset: int = 0, scale: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """rows [Bq,4] = {lse_neg, n_neg, diag, lse_all} (fp32), scal [8] (fp64) — see mi_b200.h."""
    _need_cuda(Q, K, sid_q, sid_k)
    lib = _lib.load()
    Q, K = Q.contiguous(), K.contiguous()
    sid_q, sid_k = sid_q.to(torch.int32).contiguous(), sid_k.to(torch.int32).contiguous()
    Bq, D = Q.shape
    Bk = K.shape[0]
    rows = torch.empty((Bq, 4), dtype=torch.float32, device=Q.device)
    scal = torch.empty(8, dtype=torch.float64, device=Q.device)
    nbytes = lib.mi_score_stats_workspace_bytes(Bq, Bk, D)
    ws = workspace(nbytes, Q.device)
    _check(lib.mi_score_stats(_ptr(Q), D, _ptr(K), D, _ptr(sid_q), _ptr(sid_k), q_offset, Bq, Bk, D, scale,
                              _ptr(rows), _ptr(scal), _ptr(ws), ws.numel(), _stream()), "mi_score_stats")
    return rows, scal


def score_grad(Q, K, sid_q, sid_k, q_offset: int, scale: float,
               refq: Optional[torch.Tensor], wq: float, refk: Optional[torch.Tensor], wk: float,
               include_diag: bool, precision: str, alpha: float, gamma: float, sub: Optional[torch.Tensor],
               want_f32: bool = True, want_bf16: bool = False):
    """O = alpha * (G K - gamma * sub) with G recomputed tile by tile (mi_score_grad)."""
    _need_cuda(Q, K, sid_q, sid_k, refq, refk, sub)
    lib = _lib.load()
    Q, K = Q.contiguous(), K.contiguous()
    sid_q, sid_k = sid_q.to(torch.int32).contiguous(), sid_k.to(torch.int32).contiguous()
    Bq, D = Q.shape
    Bk = K.shape[0]
    prec = PRECISION[precision]
    o32 = torch.empty((Bq, D), dtype=torch.float32, device=Q.device) if want_f32 else None
    o16 = torch.empty((Bq, D), dtype=torch.bfloat16, device=Q.device) if want_bf16 else None
    refq = None if refq is None else refq.float().contiguous()
    refk = None if refk is None else refk.float().contiguous()
    sub = None if sub is None else sub.contiguous()
    nbytes = lib.mi_score_grad_workspace_bytes(Bq, Bk, D, prec)
    ws = workspace(nbytes, Q.device)
    _check(lib.mi_score_grad(_ptr(Q), D, _ptr(K), D, _ptr(sid_q), _ptr(sid_k), q_offset, Bq, Bk, D, scale,
                             _ptr(refq), wq, _ptr(refk), wk, int(include_diag), prec, alpha, gamma,
                             _ptr(sub), D, _ptr(o32), _ptr(o16), D, _ptr(ws), ws.numel(), _stream()), "mi_score_grad")
    return o32, o16


def critic_loss_fwd_bwd(X: torch.Tensor, Y: torch.Tensor, W: Optional[torch.Tensor], sid: torch.Tensor,
                        estimator: str = "dv", precision: str = "fast", inv_tau: float = 1.0,
                        need_grads: bool = True):
    """The whole path on one GPU (mi_critic_loss_fwd_bwd).  Returns (loss_out fp64[8], dX, dY, dW)."""
    _need_cuda(X, Y, W, sid)
    lib = _lib.load()
    X, Y = as_bf16(X), as_bf16(Y)
    W = None if W is None else as_bf16(W)
    sid = sid.to(torch.int32).contiguous()
    B, D = X.shape
    critic = 1 if W is not None else 0
    est, prec = ESTIMATOR[estimator], PRECISION[precision]
    loss = torch.empty(8, dtype=torch.float64, device=X.device)
    dX = dY = dW = None
    if need_grads:
        dX = torch.empty((B, D), dtype=torch.float32, device=X.device)
        dY = torch.empty((B, D), dtype=torch.float32, device=X.device)
        if W is not None:
            dW = torch.empty((D, D), dtype=torch.float32, device=X.device)
    nbytes = lib.mi_critic_workspace_bytes(B, D, critic, est, prec, int(need_grads))
    ws = workspace(nbytes, X.device)
    _check(lib.mi_critic_loss_fwd_bwd(_ptr(X), _ptr(Y), _ptr(W), _ptr(sid), B, D, critic, est, prec, inv_tau,
                                      _ptr(loss), _ptr(dX), _ptr(dY), _ptr(dW), _ptr(ws), ws.numel(), _stream()),
           "mi_critic_loss_fwd_bwd")
    return loss, dX, dY, dW


def launch_count() -> int:
    return int(_lib.load().mi_launch_count())
