"""Thin torch-tensor wrappers over the C-ABI (include/b200nlp.h).

PyTorch is only the device-memory carrier here: every function validates shapes/dtypes, allocates outputs with
torch.empty and forwards raw pointers + the current CUDA stream to libb200nlp.so.  No op has a fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

BF16 = torch.bfloat16


def _chk(t: torch.Tensor, name: str, dtype=BF16):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


_workspaces = {}


def _workspace(nbytes: int, device, tag: str = "") -> torch.Tensor:
    """Grow-only scratch buffer per (device, tag); kernels never allocate, the caller (this module) does."""
    key = (device, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _zero_workspace(nbytes: int, device, tag: str) -> torch.Tensor:
    """Scratch buffer that is zero-filled when (re)allocated; its users must hand it back zeroed."""
    key = (device, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 1), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


# ----------------------------------------------------------------------------------------------------------
# GEMM
# ----------------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, trans_a: bool = False,
         trans_b: bool = False, accumulate: bool = False, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, cta_group: int = 2, max_ctas: int = 0) -> torch.Tensor:
    """out (+)= op(a) @ op(b) (+ bias)   or   out = bf16(bf16(op(a) @ op(b) + bias) + residual).
    a, b 2-D bf16 with unit inner stride.
    trans_a: a is stored [K, M];  trans_b: b is stored [N, K].  Default b layout [K, N] is Paddle's nn.Linear weight."""
    _chk(a, "a"); _chk(b, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if trans_a:
        K, M = a.shape
    else:
        M, K = a.shape
    if trans_b:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    if K != Kb:
        raise ValueError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        if accumulate:
            raise ValueError("gemm: accumulate=True needs an output tensor")
        out = torch.empty(M, N, dtype=BF16, device=a.device)
    _chk(out, "out")
    assert out.shape == (M, N) and out.stride(1) == 1
    if bias is not None:
        _chk(bias, "bias", torch.float32)
        assert bias.numel() == N
    ldr = 0
    if residual is not None:
        _chk(residual, "residual")
        assert residual.shape == (M, N) and residual.stride(1) == 1 and not accumulate
        ldr = residual.stride(0)
    call("b200_gemm_bf16_ex", ptr(a), ptr(b), ptr(out), ptr(bias), ptr(residual), M, N, K, a.stride(0), b.stride(0),
         out.stride(0), ldr, 1 if trans_a else 0, 0 if trans_b else 1, 1 if accumulate else 0, cta_group, max_ctas,
         stream_ptr())
    return out


def gemm_swiglu(x: torch.Tensor, w_gate_up: torch.Tensor, gate_up: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, cta_group: int = 2, store_gate_up: bool = True):
    """(gate_up [M, 2I], m [M, I]) = fused gate|up projection + SwiGLU (one tcgen05 GEMM; the epilogue holds gate and up of the
    same channels).  w_gate_up is the reference-layout fused weight [K, 2I] (gate | up).  Requires I % 128 == 0.
    store_gate_up=False (inference): only m is written and (None, m) is returned."""
    _chk(x, "x"); _chk(w_gate_up, "w_gate_up")
    assert x.dim() == 2 and w_gate_up.dim() == 2 and x.stride(1) == 1 and w_gate_up.stride(1) == 1
    M, K = x.shape
    Kw, two_i = w_gate_up.shape
    if K != Kw or two_i % 2:
        raise ValueError(f"gemm_swiglu: shapes {tuple(x.shape)} x {tuple(w_gate_up.shape)}")
    inter = two_i // 2
    if gate_up is None and store_gate_up:
        gate_up = torch.empty(M, two_i, dtype=BF16, device=x.device)
    if out is None:
        out = torch.empty(M, inter, dtype=BF16, device=x.device)
    call("b200_gemm_swiglu_bf16", ptr(x), ptr(w_gate_up), ptr(gate_up) if store_gate_up else 0, ptr(out), M, inter, K, x.stride(0),
         w_gate_up.stride(0), gate_up.stride(0) if store_gate_up else two_i, out.stride(0), cta_group, stream_ptr())
    return (gate_up if store_gate_up else None), out


def gemm_swiglu_bwd(dy: torch.Tensor, w_down: torch.Tensor, gate_up: torch.Tensor, dgate_up: Optional[torch.Tensor] = None,
                    cta_group: int = 2) -> torch.Tensor:
    """dgate_up [M, 2I] = SwiGLU backward of d(m) = dy @ w_down^T, computed in the GEMM epilogue (d(m) is never written).
    w_down is the reference-layout down_proj weight [I, h]; gate_up the saved projection [M, 2I].  Requires I % 64 == 0."""
    _chk(dy, "dy"); _chk(w_down, "w_down"); _chk(gate_up, "gate_up")
    M, K = dy.shape
    inter, Kw = w_down.shape
    assert K == Kw and gate_up.shape == (M, 2 * inter) and dy.stride(1) == 1 and w_down.stride(1) == 1 and gate_up.stride(1) == 1
    if dgate_up is None:
        dgate_up = torch.empty_like(gate_up)
    call("b200_gemm_swiglu_bwd_bf16", ptr(dy), ptr(w_down), ptr(gate_up), ptr(dgate_up), M, inter, K, dy.stride(0), w_down.stride(0),
         gate_up.stride(0), dgate_up.stride(0), cta_group, stream_ptr())
    return dgate_up


def gemm_skinny(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, trans_b: bool = False,
                bias: Optional[torch.Tensor] = None, split_k: int = 0) -> torch.Tensor:
    """Decode-step GEMM (few token rows, weight-streaming bound): split-K over all SMs, fp32 TMA reduce, one rounding."""
    _chk(a, "a"); _chk(b, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N, Kb = (b.shape if trans_b else (b.shape[1], b.shape[0]))
    if K != Kb:
        raise ValueError(f"gemm_skinny: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, dtype=BF16, device=a.device)
    if bias is not None:
        _chk(bias, "bias", torch.float32)
    ws = _zero_workspace(M * N * 4, a.device, "splitk")
    call("b200_gemm_bf16_splitk", ptr(a), ptr(b), ptr(out), ptr(bias), ptr(ws), M, N, K, a.stride(0), b.stride(0),
         out.stride(0), 0, 0 if trans_b else 1, split_k, stream_ptr())
    return out


def gemm_skinny_f32(a: torch.Tensor, b: torch.Tensor, *, trans_b: bool = False, split_k: int = 0, tag: str = "splitk_f32"):
    """Split-K GEMM that leaves its result as fp32 sums in a (zero-on-entry) workspace [M, N]; the consumer kernel
    (add_rmsnorm_f32 / decode_rope_append_f32) rounds once and re-zeroes it.  Returns the fp32 workspace view."""
    _chk(a, "a"); _chk(b, "b")
    M, K = a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    ws = _zero_workspace(M * N * 4, a.device, tag)
    call("b200_gemm_bf16_splitk", ptr(a), ptr(b), None, None, ptr(ws), M, N, K, a.stride(0), b.stride(0), N, 0,
         0 if trans_b else 1, split_k, stream_ptr())
    return ws[: M * N * 4].view(torch.float32).view(M, N)


# ----------------------------------------------------------------------------------------------------------
# RMSNorm
# ----------------------------------------------------------------------------------------------------------
def rmsnorm_fwd(x: torch.Tensor, w: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None,
                rstd: Optional[torch.Tensor] = None):
    _chk(x, "x"); _chk(w, "w")
    h = x.shape[-1]
    rows = x.numel() // h
    assert x.is_contiguous() and w.numel() == h
    if out is None:
        out = torch.empty_like(x)
    if rstd is None:
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    call("b200_rmsnorm_fwd", ptr(x), ptr(w), ptr(out), ptr(rstd), rows, h, float(eps), stream_ptr())
    return out, rstd


def rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, rstd: torch.Tensor, dw: torch.Tensor,
                dres: Optional[torch.Tensor] = None, accumulate_dw: bool = True, dx: Optional[torch.Tensor] = None):
    _chk(dy, "dy"); _chk(x, "x"); _chk(w, "w"); _chk(dw, "dw"); _chk(rstd, "rstd", torch.float32)
    h = x.shape[-1]
    rows = x.numel() // h
    assert dy.is_contiguous() and x.is_contiguous() and dw.numel() == h
    if dx is None:
        dx = torch.empty_like(x)
    ws = _workspace(_lib.load().b200_rmsnorm_bwd_workspace_bytes(rows, h), x.device, "rmsnorm_bwd")
    call("b200_rmsnorm_bwd", ptr(dy), ptr(x), ptr(w), ptr(rstd), ptr(dres), ptr(dx), ptr(dw), 1 if accumulate_dw else 0,
         ptr(ws), rows, h, stream_ptr())
    return dx


def colsum(a: torch.Tensor, out: torch.Tensor, accumulate: bool = True):
    """out[n] (+)= sum over rows of a[rows, n] (a may be a column slice: unit inner stride, any row stride)."""
    _chk(a, "a"); _chk(out, "out")
    rows, n = a.shape
    assert a.stride(1) == 1 and out.numel() == n
    ws = _workspace(_lib.load().b200_colsum_workspace_bytes(rows, n), a.device, "colsum")
    call("b200_colsum_bf16", ptr(a), ptr(out), 1 if accumulate else 0, ptr(ws), rows, n, a.stride(0), stream_ptr())
    return out


# ----------------------------------------------------------------------------------------------------------
# RoPE
# ----------------------------------------------------------------------------------------------------------
def rope_inv_freq(head_dim: int, theta: float, scaling: Optional[dict] = None, seq_len: int = 0,
                  max_position_embeddings: int = 0) -> torch.Tensor:
    """fp32 inverse frequencies [head_dim/2] on the CPU for the reference's rotary variants (llama/modeling.py):
      None / {}                          LlamaRotaryEmbedding                    :402-439
      {"rope_type": "llama3", factor, low_freq_factor, high_freq_factor, original_max_position_embeddings}
                                         Llama3RotaryEmbedding                   :520-554  (Llama-3.1 wavelength bands)
      {"type": "ntk", "factor": f}       LlamaNTKScalingRotaryEmbedding          :464-470  (base * f^(d/(d-2)))
      {"type": "dynamic_ntk", "factor"}  LlamaDynamicNTKScalingRotaryEmbedding   :473-517  (base rescaled only when
                                         seq_len > max_position_embeddings)
      {"type": "linear", "factor": f}    handled in rope_tables (positions / f)  :440-461"""
    kind = None if not scaling else (scaling.get("rope_type") or scaling.get("type"))
    base = float(theta)
    d = head_dim
    if kind == "ntk":
        base = base * float(scaling["factor"]) ** (d / (d - 2))
    elif kind == "dynamic_ntk" and max_position_embeddings and seq_len > max_position_embeddings:
        f = float(scaling["factor"])
        alpha = (f * seq_len / max_position_embeddings) - (f - 1)
        base = base * alpha ** (d / (d - 2))
    inv_freq = 1.0 / (base ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    if kind == "llama3":
        factor = float(scaling["factor"])
        lo, hi = float(scaling["low_freq_factor"]), float(scaling["high_freq_factor"])
        orig = float(scaling["original_max_position_embeddings"])
        low_wavelen, high_wavelen = orig / lo, orig / hi
        out = []
        for freq in inv_freq:                          # per-frequency loop in fp32 tensor arithmetic, as the reference (:540-552)
            wavelen = 2 * math.pi / freq
            if wavelen < high_wavelen:
                out.append(freq)
            elif wavelen > low_wavelen:
                out.append(freq / factor)
            else:
                smooth = (orig / wavelen - lo) / (hi - lo)
                out.append((1 - smooth) * freq / factor + smooth * freq)
        inv_freq = torch.stack(out).to(torch.float32)
    elif kind not in (None, "ntk", "dynamic_ntk", "linear", "default"):
        raise ValueError(f"Unknown RoPE scaling type {kind}")       # llama/modeling.py:864
    return inv_freq


def rope_tables(head_dim: int, max_pos: int, theta: float, device, scaling: Optional[dict] = None,
                max_position_embeddings: int = 0):
    """fp32 cos/sin tables [max_pos, head_dim/2], computed on the CPU exactly as the reference does
    (llama/modeling.py:409-423) so that host and device share bits, then uploaded once."""
    inv_freq = rope_inv_freq(head_dim, theta, scaling, seq_len=max_pos, max_position_embeddings=max_position_embeddings)
    t = torch.arange(max_pos, dtype=torch.float32)
    if scaling and (scaling.get("rope_type") or scaling.get("type")) == "linear":
        t = t / float(scaling["factor"])
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    return freqs.cos().contiguous().to(device), freqs.sin().contiguous().to(device)


def rope_inplace(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, seq_len: int, num_heads: int, head_dim: int,
                 position_ids: Optional[torch.Tensor] = None, backward: bool = False):
    """x: [tokens, ld] view whose first num_heads*head_dim columns are the heads to rotate (row stride = ld)."""
    _chk(x, "x"); _chk(cos, "cos", torch.float32); _chk(sin, "sin", torch.float32)
    assert x.dim() == 2 and x.stride(1) == 1
    tokens = x.shape[0]
    if position_ids is not None:
        _chk(position_ids, "position_ids", torch.int32)
        assert position_ids.numel() == tokens
    else:
        assert cos.shape[0] >= seq_len
    call("b200_rope_inplace", ptr(x), ptr(cos), ptr(sin), ptr(position_ids), tokens, seq_len, x.stride(0), num_heads,
         head_dim, 1 if backward else 0, stream_ptr())
    return x


# ----------------------------------------------------------------------------------------------------------
# SwiGLU / embedding
# ----------------------------------------------------------------------------------------------------------
def swiglu_fwd(gate_up: torch.Tensor, out: Optional[torch.Tensor] = None):
    _chk(gate_up, "gate_up")
    rows, two_i = gate_up.shape
    assert gate_up.is_contiguous() and two_i % 2 == 0
    inter = two_i // 2
    if out is None:
        out = torch.empty(rows, inter, dtype=BF16, device=gate_up.device)
    call("b200_swiglu_fwd", ptr(gate_up), ptr(out), rows, inter, stream_ptr())
    return out


def swiglu_fwd_f32(acc_f32: torch.Tensor, out: Optional[torch.Tensor] = None):
    """SwiGLU fed by the fp32 split-K workspace [rows, 2I] of the ffn1 GEMM (rounded here, workspace re-zeroed)."""
    _chk(acc_f32, "acc_f32", torch.float32)
    rows, two_i = acc_f32.shape
    assert acc_f32.is_contiguous() and two_i % 2 == 0
    inter = two_i // 2
    if out is None:
        out = torch.empty(rows, inter, dtype=BF16, device=acc_f32.device)
    call("b200_swiglu_fwd_f32", ptr(acc_f32), ptr(out), rows, inter, stream_ptr())
    return out


def gemm_swiglu_skinny(x: torch.Tensor, w_gate_up: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decode-step ffn1 + SwiGLU (M <= 64 token rows) in one weight-streaming kernel (two CTAs per SM, swapped operands);
    w_gate_up is the reference-layout fused weight [K, 2I] (gate | up), I % 64 == 0."""
    _chk(x, "x"); _chk(w_gate_up, "w_gate_up")
    M, K = x.shape
    inter = w_gate_up.shape[1] // 2
    assert w_gate_up.shape[0] == K and x.stride(1) == 1 and w_gate_up.stride(1) == 1
    if out is None:
        out = torch.empty(M, inter, dtype=BF16, device=x.device)
    assert out.shape == (M, inter) and out.stride(1) == 1
    call("b200_gemm_swiglu_skinny", ptr(x), ptr(w_gate_up), ptr(out), M, inter, K, x.stride(0), w_gate_up.stride(0),
         out.stride(0), stream_ptr())
    return out


def decode_layer_chain(attn: torch.Tensor, w_o: torch.Tensor, w_ffn_ln: torch.Tensor, w_ffn1: torch.Tensor, w_ffn2: torch.Tensor,
                       w_next_ln: Optional[torch.Tensor], w_next_qkv: Optional[torch.Tensor], residual: torch.Tensor, eps: float,
                       qkv_tag: str = "splitk_qkv"):
    """One persistent kernel for the GEMM chain of a decode layer (M <= 64 rows): out-linear, residual + ffn RMSNorm, ffn1 +
    SwiGLU, ffn2, residual + the next layer's RMSNorm, the next layer's QKV projection (see b200_decode_layer_chain).
    `residual` [M, h] is updated IN PLACE.  Returns the fp32 QKV accumulation [M, qkv_n] of the next layer (the workspace
    decode_rope_append_f32 consumes and re-zeroes) or None for the last layer (w_next_qkv None)."""
    for t, n in ((attn, "attn"), (w_o, "w_o"), (w_ffn_ln, "w_ffn_ln"), (w_ffn1, "w_ffn1"), (w_ffn2, "w_ffn2"), (residual, "residual")):
        _chk(t, n)
    M, aw = attn.shape
    h = w_o.shape[1]
    inter = w_ffn2.shape[0]
    assert w_o.shape == (aw, h) and w_ffn1.shape == (h, 2 * inter) and w_ffn2.shape == (inter, h) and residual.shape == (M, h)
    assert all(t.is_contiguous() for t in (attn, w_o, w_ffn1, w_ffn2, residual))
    dev = attn.device
    ln_buf = _workspace(M * h * 2, dev, "chain_ln")
    act_buf = _workspace(M * inter * 2, dev, "chain_act")
    acc_h = _zero_workspace(M * h * 4, dev, "chain_h")
    sync = _zero_workspace(_lib.load().b200_decode_layer_chain_workspace_bytes(), dev, "chain_sync")
    acc_qkv, qkv_n = None, 0
    if w_next_qkv is not None:
        _chk(w_next_qkv, "w_next_qkv"); _chk(w_next_ln, "w_next_ln")
        qkv_n = w_next_qkv.shape[0]
        assert w_next_qkv.shape == (qkv_n, h) and w_next_qkv.is_contiguous()
        acc_qkv = _zero_workspace(M * qkv_n * 4, dev, qkv_tag)
    call("b200_decode_layer_chain", ptr(attn), ptr(w_o), ptr(w_ffn_ln), ptr(w_ffn1), ptr(w_ffn2), ptr(w_next_ln), ptr(w_next_qkv),
         ptr(residual), ptr(ln_buf), ptr(act_buf), ptr(acc_h), ptr(acc_qkv), ptr(sync), M, h, aw, inter, qkv_n, float(eps),
         stream_ptr())
    if acc_qkv is None:
        return None
    return acc_qkv[: M * qkv_n * 4].view(torch.float32).view(M, qkv_n)


def swiglu_bwd(gate_up: torch.Tensor, dout: torch.Tensor, dgate_up: Optional[torch.Tensor] = None):
    _chk(gate_up, "gate_up"); _chk(dout, "dout")
    rows, two_i = gate_up.shape
    inter = two_i // 2
    assert gate_up.is_contiguous() and dout.is_contiguous() and dout.shape == (rows, inter)
    if dgate_up is None:
        dgate_up = torch.empty_like(gate_up)
    call("b200_swiglu_bwd", ptr(gate_up), ptr(dout), ptr(dgate_up), rows, inter, stream_ptr())
    return dgate_up


def embedding_fwd(ids: torch.Tensor, table: torch.Tensor, out: Optional[torch.Tensor] = None):
    _chk(ids, "ids", torch.int64); _chk(table, "table")
    tokens = ids.numel()
    vocab, h = table.shape
    assert ids.is_contiguous() and table.is_contiguous()
    if out is None:
        out = torch.empty(tokens, h, dtype=BF16, device=table.device)
    call("b200_embedding_fwd", ptr(ids), ptr(table), ptr(out), tokens, h, vocab, stream_ptr())
    return out


def embedding_bwd(ids: torch.Tensor, dout: torch.Tensor, dtable: torch.Tensor):
    _chk(ids, "ids", torch.int64); _chk(dout, "dout"); _chk(dtable, "dtable")
    vocab, h = dtable.shape
    assert dout.is_contiguous() and dtable.is_contiguous()
    call("b200_embedding_bwd", ptr(ids), ptr(dout), ptr(dtable), ids.numel(), h, vocab, stream_ptr())
    return dtable


# ----------------------------------------------------------------------------------------------------------
# Flash attention
# ----------------------------------------------------------------------------------------------------------
def _tok_stride(t: torch.Tensor, heads: int, d: int) -> int:
    """t: [B, S, heads, d] view with unit stride in d, d-stride between heads and a uniform token stride."""
    B, S, H, D = t.shape
    assert (H, D) == (heads, d) and t.stride(3) == 1 and (H == 1 or t.stride(2) == d)
    ld = t.stride(1)
    assert B == 1 or t.stride(0) == S * ld, "batch stride must be S * token stride"
    return ld


def _mask_rows(mask_start, B, S):
    if mask_start is None:
        return None
    _chk(mask_start, "mask_start", torch.int32)
    assert mask_start.is_contiguous() and mask_start.numel() == B * S, "mask_start must be int32 [B, S]"
    return mask_start


def flash_attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, softmax_scale: Optional[float] = None,
                   out: Optional[torch.Tensor] = None, mask_start: Optional[torch.Tensor] = None):
    """Causal GQA attention.  q [B,S,nh,128], k/v [B,S,kvh,128] (may be strided views of a packed QKV buffer).
    mask_start [B,S] int32 (optional): FlashMask start rows — row i sees column c iff c <= i < mask_start[b, c].
    Returns (o [B,S,nh,128] contiguous, lse [B,nh,S] fp32)."""
    _chk(q, "q"); _chk(k, "k"); _chk(v, "v")
    B, S, nh, d = q.shape
    kvh = k.shape[2]
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(d)
    if out is None:
        out = torch.empty(B, S, nh, d, dtype=BF16, device=q.device)
    lse = torch.empty(B, nh, S, dtype=torch.float32, device=q.device)
    call("b200_fa_fwd_flashmask", ptr(q), ptr(k), ptr(v), ptr(out), ptr(lse), ptr(_mask_rows(mask_start, B, S)), B, S, nh, kvh,
         d, _tok_stride(q, nh, d), _tok_stride(k, kvh, d), _tok_stride(v, kvh, d), _tok_stride(out, nh, d),
         float(softmax_scale), stream_ptr())
    return out, lse


def flash_attn_bwd(q, k, v, o, dout, lse, dq, dk, dv, softmax_scale: Optional[float] = None, mask_start=None):
    """Gradients written into dq/dk/dv (views allowed, e.g. slices of a packed dQKV buffer)."""
    for name, t in (("q", q), ("k", k), ("v", v), ("o", o), ("dout", dout), ("dq", dq), ("dk", dk), ("dv", dv)):
        _chk(t, name)
    _chk(lse, "lse", torch.float32)
    B, S, nh, d = q.shape
    kvh = k.shape[2]
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(d)
    ws = _workspace(_lib.load().b200_fa_bwd_workspace_bytes(B, S, nh, d), q.device, "fa_bwd")
    call("b200_fa_bwd_flashmask", ptr(q), ptr(k), ptr(v), ptr(o), ptr(dout), ptr(lse), ptr(_mask_rows(mask_start, B, S)),
         ptr(dq), ptr(dk), ptr(dv), ptr(ws), B, S, nh, kvh, d, _tok_stride(q, nh, d), _tok_stride(k, kvh, d), _tok_stride(v, kvh, d), _tok_stride(o, nh, d),
         _tok_stride(dout, nh, d), _tok_stride(dq, nh, d), _tok_stride(dk, kvh, d), _tok_stride(dv, kvh, d),
         float(softmax_scale), stream_ptr())
    return dq, dk, dv


# ----------------------------------------------------------------------------------------------------------
# Criterion / sampling
# ----------------------------------------------------------------------------------------------------------
def ce_fwd(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100):
    """Returns (loss_out [2] = (masked mean loss, count), loss_tok [T], lse [T])."""
    _chk(logits, "logits"); _chk(labels, "labels", torch.int64)
    T, V = logits.shape
    assert logits.stride(1) == 1 and labels.numel() == T and labels.is_contiguous()
    loss_tok = torch.empty(T, dtype=torch.float32, device=logits.device)
    lse = torch.empty(T, dtype=torch.float32, device=logits.device)
    loss_out = torch.empty(2, dtype=torch.float32, device=logits.device)
    call("b200_ce_fwd", ptr(logits), ptr(labels), ptr(loss_tok), ptr(lse), ptr(loss_out), T, V, logits.stride(0),
         ignore_index, stream_ptr())
    return loss_out, loss_tok, lse


def ce_bwd_(logits: torch.Tensor, labels: torch.Tensor, loss_tok, lse, loss_out, grad_scale: float = 1.0,
            grad_scale_dev: Optional[torch.Tensor] = None):
    """Overwrites logits with dlogits.  grad_scale_dev: optional fp32 device scalar multiplied into grad_scale."""
    T, V = logits.shape
    if grad_scale_dev is not None:
        _chk(grad_scale_dev, "grad_scale_dev", torch.float32)
    call("b200_ce_bwd", ptr(logits), ptr(labels), ptr(loss_tok), ptr(lse), ptr(loss_out), float(grad_scale),
         ptr(grad_scale_dev), T, V, logits.stride(0), stream_ptr())
    return logits


def argmax(logits: torch.Tensor) -> torch.Tensor:
    _chk(logits, "logits")
    rows, V = logits.shape
    out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    call("b200_argmax_bf16", ptr(logits), ptr(out), rows, V, logits.stride(0), stream_ptr())
    return out


# ----------------------------------------------------------------------------------------------------------
# Optimizer
# ----------------------------------------------------------------------------------------------------------
def grad_sqnorm(grads: torch.Tensor, scale: float = 1.0, out: Optional[torch.Tensor] = None):
    _chk(grads, "grads")
    assert grads.is_contiguous()
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=grads.device)
    ws = _workspace(_lib.load().b200_grad_sqnorm_workspace_bytes(), grads.device, "sqnorm")
    call("b200_grad_sqnorm", ptr(grads), ptr(out), ptr(ws), grads.numel(), float(scale), stream_ptr())
    return out


def adamw_step(params, grads, master, exp_avg, exp_avg_sq, sqnorm, *, decay_end: int, lr: float, beta1: float,
               beta2: float, eps: float, weight_decay: float, step: int, grad_scale: float = 1.0,
               max_grad_norm: float = 1.0):
    _chk(params, "params"); _chk(grads, "grads")
    for n_, t in (("master", master), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _chk(t, n_, torch.float32)
    n = params.numel()
    call("b200_adamw_step", ptr(params), ptr(grads), ptr(master), ptr(exp_avg), ptr(exp_avg_sq), ptr(sqnorm), n,
         decay_end, float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale),
         float(max_grad_norm), stream_ptr())


def bf16_to_f32(src: torch.Tensor, dst: torch.Tensor):
    _chk(src, "src"); _chk(dst, "dst", torch.float32)
    call("b200_bf16_to_f32", ptr(src), ptr(dst), src.numel(), stream_ptr())
    return dst


# ----------------------------------------------------------------------------------------------------------
# Generation path
# ----------------------------------------------------------------------------------------------------------
def add_rmsnorm(x, residual, w, eps, want_normed=True, want_residual=True):
    """(normed, residual_out) = fused_rms_norm(x, w, residual=residual); either output may be skipped."""
    _chk(x, "x")
    h = x.shape[-1]
    rows = x.numel() // h
    normed = torch.empty_like(x) if want_normed else None
    res_out = torch.empty_like(x) if want_residual else None
    call("b200_add_rmsnorm", ptr(x), ptr(residual), ptr(w), ptr(normed), ptr(res_out), rows, h, float(eps), stream_ptr())
    return normed, res_out


def add_rmsnorm_f32(x_f32, residual, w, eps, want_normed=True, want_residual=True):
    """add_rmsnorm whose x is the fp32 split-K workspace of the producing GEMM (consumed and re-zeroed)."""
    _chk(x_f32, "x_f32", torch.float32)
    rows, h = x_f32.shape
    normed = torch.empty(rows, h, dtype=BF16, device=x_f32.device) if want_normed else None
    res_out = torch.empty(rows, h, dtype=BF16, device=x_f32.device) if want_residual else None
    call("b200_add_rmsnorm_f32", ptr(x_f32), ptr(residual), ptr(w), ptr(normed), ptr(res_out), rows, h, float(eps), stream_ptr())
    return normed, res_out


def decode_rope_append_f32(acc_f32, bias, cache, cos, sin, seq_lens, nh, kvh, d):
    """RoPE + cache append on the fp32 split-K QKV accumulation; returns the bf16 packed projection [B, (nh+2kvh)*d]."""
    _chk(acc_f32, "acc_f32", torch.float32); _chk(cache, "cache"); _chk(seq_lens, "seq_lens", torch.int32)
    B, n = acc_f32.shape
    qkv = torch.empty(B, n, dtype=BF16, device=acc_f32.device)
    call("b200_decode_rope_append_f32", ptr(qkv), ptr(acc_f32), ptr(bias), ptr(cache), ptr(cos), ptr(sin), ptr(seq_lens), B,
         nh, kvh, d, cache.shape[3], n, stream_ptr())
    return qkv


def write_cache_kv(qkv, cache, seq_lens, B, S, nh, kvh, d):
    """qkv [B*S, ld] (post-RoPE) -> cache [2, B, kvh, max_len, d] for s < seq_lens[b]."""
    _chk(qkv, "qkv"); _chk(cache, "cache")
    assert cache.is_contiguous() and cache.shape[0] == 2 and cache.shape[1] == B and cache.shape[2] == kvh
    max_len = cache.shape[3]
    call("b200_write_cache_kv", ptr(qkv), ptr(cache), ptr(seq_lens), B, S, nh, kvh, d, max_len, qkv.stride(0), stream_ptr())


def decode_rope_append(qkv, cache, cos, sin, seq_lens, nh, kvh, d):
    _chk(qkv, "qkv"); _chk(cache, "cache"); _chk(seq_lens, "seq_lens", torch.int32)
    B = qkv.shape[0]
    call("b200_decode_rope_append", ptr(qkv), ptr(cache), ptr(cos), ptr(sin), ptr(seq_lens), B, nh, kvh, d, cache.shape[3],
         qkv.stride(0), stream_ptr())


def decode_attention(qkv, cache, seq_lens, nh, kvh, d, softmax_scale=None, out=None, num_splits: int = 0, impl: str = "tc"):
    """impl "tc": persistent tcgen05 kernel (TMA-streamed cache); "simt": the CUDA-core kernel (half-warp per cache row)."""
    _chk(qkv, "qkv"); _chk(cache, "cache"); _chk(seq_lens, "seq_lens", torch.int32)
    B = qkv.shape[0]
    if out is None:
        out = torch.empty(B, nh * d, dtype=BF16, device=qkv.device)
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(d)
    max_len = cache.shape[3]
    if impl == "tc":
        if num_splits <= 0:
            # persistent CTAs walk (split, b, kv head) items round-robin; split only when there are too few (b, kv head)
            # pairs to give every SM ~3 items (a split costs partial traffic, a merge launch and an item boundary)
            num_splits = max(1, min((max_len + 127) // 128, (3 * 148 + B * kvh - 1) // (B * kvh)))
        fn = "b200_decode_attention_tc"
    elif impl == "simt":
        if num_splits <= 0:
            # enough CTAs for ~8 per SM without making the ranges shorter than ~64 cache rows at full length
            num_splits = max(1, min(max_len // 64, (8 * 148 + B * kvh - 1) // (B * kvh)))
        fn = "b200_decode_attention"
    else:
        raise ValueError(f"decode_attention impl {impl!r}")
    ws = None
    if num_splits > 1:
        ws = _workspace(_lib.load().b200_decode_attention_workspace_bytes(B, nh, num_splits), qkv.device, "decode_attn")
    call(fn, ptr(qkv), ptr(cache), ptr(seq_lens), ptr(out), ptr(ws), B, nh, kvh, d, max_len,
         qkv.stride(0), float(softmax_scale), num_splits, stream_ptr())
    return out


# ---- paged ("block") KV cache: FusedBlockMultiTransformer / append_attention ----------------------------------------
def _paged_geom(key_cache, value_cache, block_tables):
    _chk(key_cache, "key_cache"); _chk(value_cache, "value_cache"); _chk(block_tables, "block_tables", torch.int32)
    assert key_cache.shape == value_cache.shape and key_cache.is_contiguous() and value_cache.is_contiguous()
    assert block_tables.dim() == 2 and block_tables.is_contiguous()
    num_blocks, kvh, block_size, d = key_cache.shape
    return num_blocks, kvh, block_size, d, block_tables.shape[1]


def write_cache_kv_paged(qkv, key_cache, value_cache, block_tables, seq_lens, B, S, nh):
    nb, kvh, bs, d, mb = _paged_geom(key_cache, value_cache, block_tables)
    call("b200_write_cache_kv_paged", ptr(qkv), ptr(key_cache), ptr(value_cache), ptr(block_tables), ptr(seq_lens), B, S, nh, kvh,
         d, bs, mb, qkv.stride(0), stream_ptr())


def decode_rope_append_paged(qkv, key_cache, value_cache, block_tables, cos, sin, seq_lens, nh, acc_f32=None, bias=None):
    """RoPE on the new token's q, k + append k, v at position seq_lens[b] of sequence b's block list.  With acc_f32 the packed
    projection arrives as the fp32 split-K workspace (rounded here, workspace re-zeroed) and `qkv` is created."""
    nb, kvh, bs, d, mb = _paged_geom(key_cache, value_cache, block_tables)
    if acc_f32 is not None:
        B = acc_f32.shape[0]
        qkv = torch.empty(B, (nh + 2 * kvh) * d, dtype=BF16, device=acc_f32.device)
    B = qkv.shape[0]
    call("b200_decode_rope_append_paged", ptr(qkv), ptr(acc_f32), ptr(bias), ptr(key_cache), ptr(value_cache), ptr(block_tables),
         ptr(cos), ptr(sin), ptr(seq_lens), B, nh, kvh, d, bs, mb, qkv.stride(0), stream_ptr())
    return qkv


def decode_attention_paged(qkv, key_cache, value_cache, block_tables, seq_lens, nh, softmax_scale=None, out=None,
                           num_splits: int = 0):
    _chk(qkv, "qkv"); _chk(seq_lens, "seq_lens", torch.int32)
    nb, kvh, bs, d, mb = _paged_geom(key_cache, value_cache, block_tables)
    B = qkv.shape[0]
    if out is None:
        out = torch.empty(B, nh * d, dtype=BF16, device=qkv.device)
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(d)
    max_len = mb * bs
    if num_splits <= 0:
        num_splits = max(1, min((max_len + 127) // 128, (3 * 148 + B * kvh - 1) // (B * kvh)))
    ws = None
    if num_splits > 1:
        ws = _workspace(_lib.load().b200_decode_attention_workspace_bytes(B, nh, num_splits), qkv.device, "decode_attn")
    call("b200_decode_attention_paged", ptr(qkv), ptr(key_cache), ptr(value_cache), ptr(block_tables), ptr(seq_lens), ptr(out),
         ptr(ws), B, nh, kvh, d, nb, bs, mb, qkv.stride(0), float(softmax_scale), num_splits, stream_ptr())
    return out


def fused_get_rotary_embedding(input_ids, position_ids, head_dim_shape_tensor, prompt_num: int = 0, theta: float = 10000.0,
                               use_neox: bool = True) -> torch.Tensor:
    """fused_get_rotary_embedding op of the reference (csrc/gpu/fused_get_rope.cu:159-223; same positional arguments):
    input_ids [bsz, seq] (only its shape is used), position_ids int64 [bsz, >= seq + prompt_num], head_dim_shape_tensor: any
    tensor whose FIRST dimension is head_dim (the reference's shape-carrier trick) or an int.
    Returns fp32 [2, bsz, 1, seq, head_dim] (cos, sin)."""
    _chk(position_ids, "position_ids", torch.int64)
    bsz, seq = input_ids.shape[0], input_ids.shape[1]
    head_dim = int(head_dim_shape_tensor) if isinstance(head_dim_shape_tensor, int) else int(head_dim_shape_tensor.shape[0])
    assert position_ids.dim() == 2 and position_ids.is_contiguous() and position_ids.shape[0] == bsz
    out = torch.empty(2, bsz, 1, seq, head_dim, dtype=torch.float32, device=position_ids.device)
    call("b200_fused_get_rotary_embedding", ptr(position_ids), ptr(out), bsz, seq, position_ids.shape[1], head_dim,
         int(prompt_num), float(theta), 1 if use_neox else 0, stream_ptr())
    return out


def append_attention(qkv, key_cache, value_cache, seq_lens_encoder, seq_lens_decoder, seq_lens_this_time, cu_seqlens_q,
                     block_tables, cos, sin, nh: int, max_q_len: int, softmax_scale=None, out=None, num_splits: int = 0):
    """append_attention of the reference (csrc/gpu/append_attention.cu:428-851) for a mixed batch over the paged cache: RoPE +
    cache append for every new token row of the packed projection qkv [token_num, (nh + 2 kvh) d] (modified in place), then
    attention of every row over its sequence's pages — prompts / prompt chunks and decode rows in one call.
    Returns out [token_num, nh * d]."""
    _chk(qkv, "qkv"); _chk(cos, "cos", torch.float32); _chk(sin, "sin", torch.float32)
    for name, t in (("seq_lens_encoder", seq_lens_encoder), ("seq_lens_decoder", seq_lens_decoder),
                    ("seq_lens_this_time", seq_lens_this_time), ("cu_seqlens_q", cu_seqlens_q)):
        _chk(t, name, torch.int32)
        assert t.is_contiguous()
    nb, kvh, bs, d, mb = _paged_geom(key_cache, value_cache, block_tables)
    token_num, ld = qkv.shape
    B = seq_lens_this_time.numel()
    assert qkv.stride(1) == 1 and ld == (nh + 2 * kvh) * d and block_tables.shape[0] == B and cu_seqlens_q.numel() >= B
    if out is None:
        out = torch.empty(token_num, nh * d, dtype=BF16, device=qkv.device)
    if softmax_scale is None:
        softmax_scale = 1.0 / math.sqrt(d)
    if num_splits <= 0:
        num_splits = max(1, min((mb * bs + 127) // 128, (3 * 148 + B * kvh - 1) // (B * kvh)))
    ws = _workspace(_lib.load().b200_append_attention_workspace_bytes(B, nh, kvh, d, num_splits), qkv.device, "append_attn")
    call("b200_append_attention", ptr(qkv), ptr(key_cache), ptr(value_cache), ptr(seq_lens_encoder), ptr(seq_lens_decoder),
         ptr(seq_lens_this_time), ptr(cu_seqlens_q), ptr(block_tables), ptr(cos), ptr(sin), ptr(out), ptr(ws), B, token_num,
         int(max_q_len), nh, kvh, d, nb, bs, mb, cos.shape[0], qkv.stride(0), out.stride(0), float(softmax_scale), num_splits,
         stream_ptr())
    return out


def get_padding_offset(input_ids, cum_offsets, token_num, seq_lens):
    """get_padding_offset_v2: returns (x_remove_padding, cum_offsets_out, padding_offset, cu_seqlens_q, cu_seqlens_k)."""
    bsz, max_len = input_ids.shape
    dev = input_ids.device
    xr = torch.zeros(int(token_num), dtype=torch.int64, device=dev)
    po = torch.zeros(int(token_num), dtype=torch.int32, device=dev)
    co = torch.zeros(bsz, dtype=torch.int32, device=dev)
    cq = torch.zeros(bsz + 1, dtype=torch.int32, device=dev)
    ck = torch.zeros(bsz + 1, dtype=torch.int32, device=dev)
    call("b200_get_padding_offset", ptr(input_ids), ptr(cum_offsets), ptr(seq_lens), ptr(xr), ptr(po), ptr(co), ptr(cq), ptr(ck),
         bsz, max_len, stream_ptr())
    return xr, co, po, cq, ck


def rebuild_padding(tmp_out, cum_offsets, seq_lens_decoder, seq_lens_encoder, max_len):
    _chk(tmp_out, "tmp_out")
    bsz, dim = seq_lens_encoder.numel(), tmp_out.shape[1]
    out = torch.zeros(bsz, dim, dtype=BF16, device=tmp_out.device)
    call("b200_rebuild_padding", ptr(tmp_out), ptr(cum_offsets), ptr(seq_lens_decoder), ptr(seq_lens_encoder), ptr(out), bsz,
         max_len, dim, stream_ptr())
    return out


def set_value_by_flags_and_idx(pre_ids_all, pre_ids_now, step_idx, stop_flags):
    bs, length = pre_ids_all.shape
    call("b200_set_value_by_flags_and_idx", ptr(stop_flags), ptr(pre_ids_all), ptr(pre_ids_now), ptr(step_idx), bs, length,
         stream_ptr())


def set_value_by_flags_and_idx_v2(pre_ids_all, input_ids, seq_lens_this_time, seq_lens_encoder, seq_lens_decoder, step_idx,
                                  stop_flags):
    bs, length = pre_ids_all.shape
    call("b200_set_value_by_flags_and_idx_v2", ptr(stop_flags), ptr(pre_ids_all), ptr(input_ids), ptr(seq_lens_encoder),
         ptr(seq_lens_decoder), ptr(step_idx), bs, length, input_ids.shape[1], stream_ptr())


def token_penalty_multi_scores(pre_ids, logits, penalty_scores, frequency_scores, presence_scores, temperatures, bad_tokens,
                               cur_len, min_len, eos_token_id):
    """In place on fp32 logits (get_token_penalty_multi_scores_v2; pass temperatures/bad_tokens=None for the v1 op)."""
    _chk(logits, "logits", torch.float32)
    bs, length = logits.shape
    ws = _workspace(bs * length * 4, logits.device, "penalty")
    call("b200_token_penalty_multi_scores", ptr(pre_ids), ptr(logits), ptr(penalty_scores), ptr(frequency_scores),
         ptr(presence_scores), ptr(temperatures), ptr(bad_tokens), ptr(cur_len), ptr(min_len), ptr(eos_token_id), ptr(ws), bs,
         length, pre_ids.shape[1], 0 if bad_tokens is None else bad_tokens.numel(), eos_token_id.numel(), stream_ptr())
    return logits


def set_stop_value_multi_ends(topk_ids, stop_flags, end_ids, seq_lens=None, next_tokens=None):
    """v2 when seq_lens / next_tokens are given, else the v1 op in mode 2.  In place."""
    v2 = seq_lens is not None
    call("b200_set_stop_value_multi_ends", ptr(stop_flags), ptr(topk_ids), ptr(next_tokens), ptr(end_ids), ptr(seq_lens),
         topk_ids.numel(), end_ids.numel(), 1 if v2 else 0, stream_ptr())


def update_inputs(stop_flags, not_need_stop, seq_lens_this_time, seq_lens_encoder, seq_lens_decoder, input_ids, stop_nums,
                  next_tokens, is_block_step):
    call("b200_update_inputs", ptr(not_need_stop), ptr(seq_lens_this_time), ptr(seq_lens_encoder), ptr(seq_lens_decoder),
         ptr(input_ids), ptr(stop_nums), ptr(stop_flags), ptr(is_block_step), ptr(next_tokens), seq_lens_this_time.numel(),
         stop_flags.numel(), input_ids.shape[1], stream_ptr())


def step_paddle(stop_flags, seq_lens_this_time, ori_seq_lens_encoder, seq_lens_encoder, seq_lens_decoder, block_tables,
                encoder_block_lens, is_block_step, step_block_list, step_lens, recover_block_list, recover_lens, need_block_list,
                need_block_len, used_list_len, free_list, free_list_len, input_ids, pre_ids, step_idx, next_tokens, block_size: int,
                encoder_decoder_block_num: int = 0, first_token_id: int = 0):
    """step_paddle(...) of the reference (csrc/gpu/step.cu:216-283; same argument order and in-place semantics)."""
    for name, t, dt in (("stop_flags", stop_flags, torch.bool), ("is_block_step", is_block_step, torch.bool),
                        ("seq_lens_this_time", seq_lens_this_time, torch.int32), ("block_tables", block_tables, torch.int32),
                        ("free_list", free_list, torch.int32), ("input_ids", input_ids, torch.int64), ("pre_ids", pre_ids, torch.int64),
                        ("step_idx", step_idx, torch.int64), ("next_tokens", next_tokens, torch.int64)):
        _chk(t, name, dt)
        assert t.is_contiguous(), name
    bsz = seq_lens_this_time.shape[0]
    call("b200_step_paddle", ptr(stop_flags), ptr(seq_lens_this_time), ptr(ori_seq_lens_encoder), ptr(seq_lens_encoder),
         ptr(seq_lens_decoder), ptr(block_tables), ptr(encoder_block_lens), ptr(is_block_step), ptr(step_block_list), ptr(step_lens),
         ptr(recover_block_list), ptr(recover_lens), ptr(need_block_list), ptr(need_block_len), ptr(used_list_len), ptr(free_list),
         ptr(free_list_len), ptr(input_ids), ptr(pre_ids), ptr(step_idx), ptr(next_tokens), bsz, int(block_size),
         block_tables.shape[1], input_ids.shape[1], pre_ids.shape[1], int(first_token_id), stream_ptr())


def generate_step_update(next_tokens, stop_flags, step_idx, max_dec_len, seq_len_decoder, pre_ids, eos_ids, out_tokens,
                         stop_count, out_col=0, out_col_dev=None):
    bs = next_tokens.numel()
    call("b200_generate_step_update", ptr(next_tokens), ptr(stop_flags), ptr(step_idx), ptr(max_dec_len), ptr(seq_len_decoder),
         ptr(pre_ids), pre_ids.shape[1], ptr(eos_ids), eos_ids.numel(), ptr(out_tokens),
         0 if out_tokens is None else out_tokens.shape[1], int(out_col), ptr(out_col_dev), ptr(stop_count), bs, stream_ptr())


def softmax_f32_(logits):
    """In-place fp32 row softmax (generation_utils.py:327)."""
    _chk(logits, "logits", torch.float32)
    rows, V = logits.shape
    assert logits.stride(1) == 1
    call("b200_softmax_f32", ptr(logits), rows, V, logits.stride(0), stream_ptr())
    return logits


def top_p_sampling_reject(probs, top_p, uniform=None, seed: int = 0, max_rounds: int = 32, generator=None):
    """top_p_sampling_reject(probs, top_p, seed) of the reference (csrc/gpu/sample_kernels/top_p_sampling_reject.cu).
    probs [bs, V] fp32, top_p [bs] fp32 -> ids [bs] int64.  `uniform` [max_rounds, bs] may be supplied (tests); otherwise it
    is drawn from `generator` (or a fresh generator seeded with `seed` when seed != 0, else torch's default CUDA generator)."""
    _chk(probs, "probs", torch.float32); _chk(top_p, "top_p", torch.float32)
    bs, V = probs.shape
    assert probs.stride(1) == 1 and top_p.numel() == bs
    if uniform is None:
        if generator is None and seed:
            generator = torch.Generator(device=probs.device)
            generator.manual_seed(int(seed))
        uniform = torch.rand(max_rounds, bs, dtype=torch.float32, device=probs.device, generator=generator)
    _chk(uniform, "uniform", torch.float32)
    assert uniform.is_contiguous() and uniform.shape == (max_rounds, bs)
    out = torch.empty(bs, dtype=torch.int64, device=probs.device)
    call("b200_top_p_sampling_reject", ptr(probs), ptr(top_p), ptr(uniform), ptr(out), bs, V, probs.stride(0), max_rounds,
         stream_ptr())
    return out


def argmax_f32(logits):
    _chk(logits, "logits", torch.float32)
    rows, V = logits.shape
    out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    call("b200_argmax_f32", ptr(logits), ptr(out), rows, V, logits.stride(0), stream_ptr())
    return out


def bf16_rows_to_f32(src, out=None):
    _chk(src, "src")
    rows, cols = src.shape
    if out is None:
        out = torch.empty(rows, cols, dtype=torch.float32, device=src.device)
    call("b200_bf16_rows_to_f32", ptr(src), ptr(out), rows, cols, src.stride(0), stream_ptr())
    return out
