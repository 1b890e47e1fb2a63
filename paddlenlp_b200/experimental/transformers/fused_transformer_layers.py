"""FusedMultiTransformer — the stacked-layer inference block of
paddlenlp/experimental/transformers/fused_transformer_layers.py (:205-345 config, :348-792 weights, :1027-1182 forward)
for the bf16, non-quantised, rmsnorm + swiglu + rotate-half-RoPE case, on the native sm_100a kernels.

Weight layouts are the reference's:
    qkv_weight    [(nh + 2*kvh) * d, h]   (transposed, trans_qkvw=True)      -> GEMM with B stored [N, K]
    linear_weight [nh * d, h]             ffn1_weight [h, 2*I] (gate | up)    ffn2_weight [I, h]
    ln_scale / ffn_ln_scale [h]
KV cache per layer: bf16 [2, B, kvh, max_len, d].
Layer loop (:1126-1174): the output norm of layer i is fused with the residual add and is layer i+1's input norm.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional

import torch

from ... import ops

BF16 = torch.bfloat16


@dataclass
class FusedMultiTransformerConfig:
    embed_dim: int
    num_heads: int
    dim_feedforward: int
    kv_num_heads: int = -1
    num_layers: int = -1
    epsilon: float = 1e-5
    norm_type: str = "rmsnorm"
    activation: str = "swiglu"
    use_neox_rotary_style: bool = True       # csrc naming: neox == rotate-half (encode_rotary_qk.cu:18-56)
    rope_theta: float = 10000.0
    max_position_embeddings: int = 4096
    qkv_bias: bool = False
    nranks: int = 1
    trans_qkvw: bool = True
    append_attn: bool = False                # FusedBlockMultiTransformer: route attention through the unified append_attention op

    def __post_init__(self):
        if self.kv_num_heads <= 0:
            self.kv_num_heads = self.num_heads
        if self.norm_type != "rmsnorm" or self.activation != "swiglu":
            raise NotImplementedError("only the rmsnorm + swiglu (Llama / Qwen2) block is implemented")
        if not self.use_neox_rotary_style:
            raise NotImplementedError("interleaved-pair RoPE is not used by Llama/Qwen2 (SURVEY.md §8 naming trap)")
        if self.nranks != 1:
            raise NotImplementedError("tensor-parallel generation is out of scope (config 5 is single-GPU)")
        if not self.trans_qkvw:
            raise NotImplementedError("trans_qkvw=False")
        if self.embed_dim // self.num_heads != 128:
            raise NotImplementedError("head_dim must be 128")


class FusedMultiTransformerBase:
    def __init__(self, config: FusedMultiTransformerConfig, device=None):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("FusedMultiTransformer needs a CUDA device: there is no CPU implementation")
            device = torch.device("cuda", torch.cuda.current_device())
        self.config = config
        self.device = torch.device(device)
        c = config
        self.h, self.nh, self.kvh, self.I, self.L = c.embed_dim, c.num_heads, c.kv_num_heads, c.dim_feedforward, c.num_layers
        self.d = self.h // self.nh
        self.qkv_n = (self.nh + 2 * self.kvh) * self.d

        def z(*shape):
            return torch.zeros(*shape, dtype=BF16, device=self.device)

        self.ln_scales = [torch.ones(self.h, dtype=BF16, device=self.device) for _ in range(self.L)]
        self.qkv_weights = [z(self.qkv_n, self.h) for _ in range(self.L)]
        self.qkv_biases: List[Optional[torch.Tensor]] = [z(self.qkv_n) if c.qkv_bias else None for _ in range(self.L)]
        self.linear_weights = [z(self.nh * self.d, self.h) for _ in range(self.L)]
        self.ffn_ln_scales = [torch.ones(self.h, dtype=BF16, device=self.device) for _ in range(self.L)]
        self.ffn1_weights = [z(self.h, 2 * self.I) for _ in range(self.L)]
        self.ffn2_weights = [z(self.I, self.h) for _ in range(self.L)]
        self._bias_f32 = [None] * self.L
        # decode-step ffn1 + SwiGLU (see the measurement note in forward()): "persistent" = 128x256-tile kernel with the SwiGLU
        # epilogue, "skinny" = swapped-operand two-CTA/SM kernel with the SwiGLU epilogue, "unfused" = GEMM + SwiGLU kernel
        self.ffn1_impl = os.environ.get("B200_DECODE_FFN1", "skinny")
        # decode step: out-linear .. next layer's qkv as ONE persistent kernel per layer (ops.decode_layer_chain) instead of six
        # programmatically chained launches.  OFF by default: measured 108 us against 89 us per layer for the six kernels
        # (tools/decode_probe.py chain, profiles/r02_decode_chain_probe.log) — the grid barriers are cheap (0.25 us) but the two
        # norm phases take 8 us each under the weight-prefetch traffic, and the chained launches already overlap their prologues
        self.layer_chain = os.environ.get("B200_DECODE_CHAIN", "0") != "0"
        self.rope = ops.rope_tables(self.d, c.max_position_embeddings, float(c.rope_theta), self.device)

    def ensure_rope(self, positions: int):
        """Grow the fp32 cos/sin tables to cover `positions` rows.  The decode kernels index them with the running sequence
        length (decode_rope_append reads row seq_len of the tables): a cache longer than config.max_position_embeddings
        must not read past the tables.  Call outside CUDA-graph capture (generate() does, before the first step)."""
        if self.rope[0].shape[0] < positions:
            self.rope = ops.rope_tables(self.d, int(positions), float(self.config.rope_theta), self.device)

    def _bias(self, i):
        if self.qkv_biases[i] is None:
            return None
        if self._bias_f32[i] is None:
            self._bias_f32[i] = self.qkv_biases[i].float()
        return self._bias_f32[i]

    def weights_changed(self):
        """Call after editing the weight tensors in place (set_state_dict does): derived copies are rebuilt lazily."""
        self._bias_f32 = [None] * self.L

    SKINNY_M = 128     # at or below this many token rows the GEMMs are weight-streaming bound: split-K kernel

    def _mm(self, a, w, trans_b=False, bias=None):
        n = w.shape[0] if trans_b else w.shape[1]
        if a.shape[0] <= self.SKINNY_M:
            if n >= 100 * 256:        # enough 256-wide column tiles to keep most SMs streaming: no split-K needed
                return ops.gemm(a, w, trans_b=trans_b, bias=bias, cta_group=1)
            return ops.gemm_skinny(a, w, trans_b=trans_b, bias=bias)
        return ops.gemm(a, w, trans_b=trans_b, bias=bias)

    # compute_qkv (:817-820): linear(ln_out, qkv_weight, transpose_weight=True)
    def compute_qkv(self, ln_out, i):
        return self._mm(ln_out, self.qkv_weights[i], trans_b=True, bias=self._bias(i))

    # ---- the three places that touch the KV cache (overridden by FusedBlockMultiTransformer) ----
    def _write_cache(self, qkv, caches, i, B, S, seq_lens_encoder, kw):
        ops.write_cache_kv(qkv, caches[i], seq_lens_encoder, B, S, self.nh, self.kvh, self.d)

    def _rope_append(self, qkv, acc, caches, i, seq_lens_decoder, kw):
        cos, sin = self.rope
        if acc is not None:
            return ops.decode_rope_append_f32(acc, self._bias(i), caches[i], cos, sin, seq_lens_decoder, self.nh, self.kvh, self.d)
        ops.decode_rope_append(qkv, caches[i], cos, sin, seq_lens_decoder, self.nh, self.kvh, self.d)
        return qkv

    def _attend(self, qkv, caches, i, seq_lens_decoder, kw):
        return ops.decode_attention(qkv, caches[i], seq_lens_decoder, self.nh, self.kvh, self.d)

    # compute_fmha (:829-882): qkv_transpose_split -> encode_rotary_qk -> write_cache_kv -> var-len attention
    def compute_fmha(self, qkv, caches, i, B, S, seq_lens_encoder, kw):
        cos, sin = self.rope
        ops.rope_inplace(qkv, cos, sin, S, self.nh + self.kvh, self.d)
        self._write_cache(qkv, caches, i, B, S, seq_lens_encoder, kw)
        q4 = qkv.view(B, S, self.qkv_n)
        qn, kn = self.nh * self.d, self.kvh * self.d
        q = q4[:, :, :qn].unflatten(2, (self.nh, self.d))
        k = q4[:, :, qn:qn + kn].unflatten(2, (self.kvh, self.d))
        v = q4[:, :, qn + kn:].unflatten(2, (self.kvh, self.d))
        attn, _ = ops.flash_attn_fwd(q, k, v)           # right-padded prompts are exact under the causal mask
        return attn.view(B * S, qn)

    # compute_mmha (:884-893): masked_multihead_attention over the cache
    def compute_mmha(self, qkv, caches, i, seq_lens_decoder, kw):
        qkv = self._rope_append(qkv, None, caches, i, seq_lens_decoder, kw)
        return self._attend(qkv, caches, i, seq_lens_decoder, kw)

    def forward(self, src: torch.Tensor, caches: List[torch.Tensor], *, B: int, S: int, seq_lens_encoder=None,
                seq_lens_decoder=None, time_step=None, **kw) -> torch.Tensor:
        """src [B*S, h] embeddings.  time_step None = prefill (S prompt positions per sequence, right padded);
        otherwise decode (S == 1, seq_lens_decoder[b] = number of cached tokens).  Returns hidden states [B*S, h]."""
        eps = self.config.epsilon
        decode = time_step is not None
        residual = src
        ln_out, _ = ops.add_rmsnorm(src, None, self.ln_scales[0], eps, want_residual=False)   # compute_layernorm_before_qkv
        fused = decode and src.shape[0] <= self.SKINNY_M and not getattr(self.config, "append_attn", False)
        if (fused and self.layer_chain and src.shape[0] <= 64 and self.I % 64 == 0 and self.h % 128 == 0 and self.h <= 8192
                and self.qkv_n % 128 == 0):
            # one persistent kernel per layer for everything between two attention calls; the residual stream lives in one buffer
            residual = src.clone()
            acc = ops.gemm_skinny_f32(ln_out, self.qkv_weights[0], trans_b=True, tag="splitk_qkv")
            for i in range(self.L):
                qkv = self._rope_append(None, acc, caches, i, seq_lens_decoder, kw)
                attn = self._attend(qkv, caches, i, seq_lens_decoder, kw)
                last = i == self.L - 1
                acc = ops.decode_layer_chain(attn, self.linear_weights[i], self.ffn_ln_scales[i], self.ffn1_weights[i],
                                             self.ffn2_weights[i], None if last else self.ln_scales[i + 1],
                                             None if last else self.qkv_weights[i + 1], residual, eps)
            return residual
        for i in range(self.L):
            if fused:
                # decode step: the split-K GEMMs leave fp32 sums that the next kernel rounds once (same rounding points,
                # three launches fewer per layer)
                acc = ops.gemm_skinny_f32(ln_out, self.qkv_weights[i], trans_b=True, tag="splitk_qkv")
                qkv = self._rope_append(None, acc, caches, i, seq_lens_decoder, kw)
                attn = self._attend(qkv, caches, i, seq_lens_decoder, kw)
                acc = ops.gemm_skinny_f32(attn, self.linear_weights[i], tag="splitk_h")
                ln_out, residual = ops.add_rmsnorm_f32(acc, residual, self.ffn_ln_scales[i], eps)
                # ffn1 + SwiGLU, measured in the 32-layer chain at context 1024 (tools/decode_ablation.py B200_FFN1=fused|epi|plain,
                # profiles/r02_decode_ablation_ffn1.log): swapped-operand kernel with the SwiGLU epilogue 4.458 ms, persistent
                # 128x256-tile kernel with the SwiGLU epilogue 4.533 ms, GEMM + SwiGLU kernel 4.630 ms.  (Both epilogues use
                # swiglu_fwd_pair: with an IEEE division + expf per element the persistent epilogue alone took 6.5 us per layer.)
                if self.ffn1_impl == "skinny" and ln_out.shape[0] <= 64 and self.I % 64 == 0:
                    # swapped-operand kernel, tile = 64 gate columns + the 64 up columns of the same channels, straight from the
                    # reference-layout weight; the activation leaves as one TMA store per tile
                    act = ops.gemm_swiglu_skinny(ln_out, self.ffn1_weights[i])
                elif self.ffn1_impl != "unfused" and self.I % 128 == 0:
                    # SwiGLU in the ffn1 epilogue of the persistent kernel: the 256-column tile pairs 128 gate columns with the
                    # 128 up columns of the same channels straight from the reference-layout weight; only the activation is stored
                    _, act = ops.gemm_swiglu(ln_out, self.ffn1_weights[i], cta_group=1, store_gate_up=False)
                else:
                    ffn1 = self._mm(ln_out, self.ffn1_weights[i])
                    act = ops.swiglu_fwd(ffn1)
                acc = ops.gemm_skinny_f32(act, self.ffn2_weights[i], tag="splitk_h")
                if i != self.L - 1:
                    ln_out, residual = ops.add_rmsnorm_f32(acc, residual, self.ln_scales[i + 1], eps)
                else:
                    _, residual = ops.add_rmsnorm_f32(acc, residual, None, eps, want_normed=False)
                continue
            qkv = self.compute_qkv(ln_out, i)
            if decode:
                attn = self.compute_mmha(qkv, caches, i, seq_lens_decoder, kw)
            else:
                attn = self.compute_fmha(qkv, caches, i, B, S, seq_lens_encoder, kw)
            out = self._mm(attn, self.linear_weights[i])                                      # compute_out_linear (:895-896)
            ln_out, residual = ops.add_rmsnorm(out, residual, self.ffn_ln_scales[i], eps)     # compute_ffn_layernorm (:937-949)
            ffn1 = self._mm(ln_out, self.ffn1_weights[i])
            act = ops.swiglu_fwd(ffn1)                                                        # fused_bias_act("swiglu") (:100-168)
            ffn2 = self._mm(act, self.ffn2_weights[i])
            if i != self.L - 1:                                                               # compute_bias_residual_layernorm (:976-999)
                ln_out, residual = ops.add_rmsnorm(ffn2, residual, self.ln_scales[i + 1], eps)
            else:
                _, residual = ops.add_rmsnorm(ffn2, residual, None, eps, want_normed=False)
        return residual

    __call__ = forward


class FusedBlockMultiTransformer(FusedMultiTransformerBase):
    """Paged ("block") KV cache variant (fused_transformer_layers.py:2192-2354, `compute_attn` -> append_attention /
    block_multihead_attention).  `caches` is the reference's list of 2*L tensors [key_cache_0, value_cache_0, key_cache_1, ...],
    each [max_block_nums, kv_num_heads, block_size, head_dim]; `block_tables` [B, max_blocks_per_seq] int32 (-1 = unused)
    arrives as a keyword argument, as in the reference.  The math is the dense path's: only cache addressing changes."""

    @staticmethod
    def _tables(kw):
        bt = kw.get("block_tables")
        if bt is None:
            raise ValueError("FusedBlockMultiTransformer needs block_tables=[B, max_blocks_per_seq] int32")
        return bt

    def _write_cache(self, qkv, caches, i, B, S, seq_lens_encoder, kw):
        ops.write_cache_kv_paged(qkv, caches[2 * i], caches[2 * i + 1], self._tables(kw), seq_lens_encoder, B, S, self.nh)

    def _rope_append(self, qkv, acc, caches, i, seq_lens_decoder, kw):
        cos, sin = self.rope
        return ops.decode_rope_append_paged(qkv, caches[2 * i], caches[2 * i + 1], self._tables(kw), cos, sin, seq_lens_decoder,
                                            self.nh, acc_f32=acc, bias=self._bias(i) if acc is not None else None)

    def _attend(self, qkv, caches, i, seq_lens_decoder, kw):
        return ops.decode_attention_paged(qkv, caches[2 * i], caches[2 * i + 1], self._tables(kw), seq_lens_decoder, self.nh)

    # ---- config.append_attn (fused_transformer_layers.py:2215-2262): ONE op does RoPE + cache append + attention for the prompt
    # rows and the decode rows alike; the padded [B, S] prefill layout is the packed layout with cu_seqlens_q[b] = b * S ----
    def _append(self, qkv, caches, i, B, S, enc, dec, this_time, kw):
        cos, sin = self.rope
        cu = torch.arange(0, (B + 1) * S, S, dtype=torch.int32, device=qkv.device)
        return ops.append_attention(qkv, caches[2 * i], caches[2 * i + 1], enc, dec, this_time, cu, self._tables(kw), cos, sin,
                                    self.nh, max_q_len=S)

    def compute_fmha(self, qkv, caches, i, B, S, seq_lens_encoder, kw):
        if not self.config.append_attn:
            return super().compute_fmha(qkv, caches, i, B, S, seq_lens_encoder, kw)
        enc = (seq_lens_encoder if seq_lens_encoder is not None
               else torch.full((B,), S, dtype=torch.int32, device=qkv.device)).to(torch.int32)
        return self._append(qkv, caches, i, B, S, enc, torch.zeros_like(enc), enc, kw)

    def compute_mmha(self, qkv, caches, i, seq_lens_decoder, kw):
        if not self.config.append_attn:
            return super().compute_mmha(qkv, caches, i, seq_lens_decoder, kw)
        B = qkv.shape[0]
        one = torch.ones(B, dtype=torch.int32, device=qkv.device)
        return self._append(qkv, caches, i, B, 1, torch.zeros_like(one), seq_lens_decoder, one, kw)
