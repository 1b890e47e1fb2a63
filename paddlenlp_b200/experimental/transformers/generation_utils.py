"""GenerationInferenceModel — the dense-KV-cache generate loop of
paddlenlp/experimental/transformers/generation_utils.py (:124-183 generate, :262-400 sample, :185-260 state update).

Per step the reference runs: set_value_by_flags_and_idx -> cast fp32 -> get_token_penalty_multi_scores -> /temperature
-> softmax -> top_p_sampling_reject -> set_stop_value_multi_ends -> save_with_output, with one host sync in the `while`
condition.  Here the whole decode step (embedding .. lm_head .. token choice .. state update) is device-resident and
replayed as a CUDA graph; the stop condition is polled every `sync_interval` steps.  Greedy decoding (top_p == 0, the
benchmark setting: predictor.py:1196-1199) takes the arg-max directly; top_p > 0 runs softmax + rejection top-p sampling
(`top_p_sampling_reject`) with uniforms from a torch CUDA generator (graph-safe).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ... import ops


class GenerationInferenceModel:
    def _prefill(self, input_ids, seq_lens_encoder, caches):
        raise NotImplementedError

    def _decode(self, tgt_ids, seq_lens_decoder, caches):
        raise NotImplementedError

    def _choose(self, logits, st):
        """fp32 cast -> penalties -> temperature -> softmax -> top-p sampling; with top_p == 0 the arg-max is taken
        directly (identical result: the sampler degenerates to top-1)."""
        greedy = st["top_p"] is None
        if greedy and st["plain"]:
            return ops.argmax(logits)
        lf = ops.bf16_rows_to_f32(logits)
        ops.token_penalty_multi_scores(st["pre_ids"], lf, st["penalty"], st["frequency"], st["presence"], st["temperature"],
                                       None, st["step_idx"], st["min_dec_len"], st["eos"])
        if greedy:
            return ops.argmax_f32(lf)
        ops.softmax_f32_(lf)
        return ops.top_p_sampling_reject(lf, st["top_p"], generator=st["generator"])

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, seq_len_encoder: Optional[torch.Tensor] = None, max_length: int = 64,
                 eos_token_id=None, cache_kvs: Optional[List[torch.Tensor]] = None, temperature: float = 1.0,
                 top_p: float = 0.0, penalty_score: float = 1.0, frequency_score: float = 0.0, presence_score: float = 0.0,
                 min_length: int = 0, use_cuda_graph: bool = True, sync_interval: int = 16, use_pdl: bool = True, seed: int = 0,
                 token_stream=None, **kwargs):
        """input_ids [B, S] (right padded); returns (ids [B, max_length], stop_flags, seq_len_decoder).
        top_p in (0, 1]: sample from the top-p nucleus (seed != 0 makes the draw reproducible); top_p 0: greedy.
        token_stream: a token_stream.TokenStream — every step's tokens are published to its pinned-host ring from inside
        the (graph-replayed) decode step, the replacement of save_with_output / save_output (generation_utils.py:353-361,
        :711-713): a reader thread sees step t while step t+1 is computed, with no host synchronisation in this loop."""
        if top_p is not None and not (0.0 <= float(top_p) <= 1.0):
            raise ValueError(f"top_p must be in [0, 1], got {top_p}")
        dev = self.device
        B, S = input_ids.shape
        ids = input_ids.to(dev, torch.int64).contiguous()
        enc = (torch.full((B,), S, dtype=torch.int32, device=dev) if seq_len_encoder is None
               else seq_len_encoder.to(dev, torch.int32).reshape(B).contiguous())
        if cache_kvs is None:
            cache_kvs = self.allocate_caches(B, S + max_length)
        if getattr(self, "block_attn", False):                   # paged cache: capacity = blocks per sequence x block size
            if self.block_tables is None or self.block_tables.shape[0] != B:
                raise ValueError("block_attn: allocate the caches with allocate_caches(batch, max_len) for this batch size")
            max_len = self.block_tables.shape[1] * cache_kvs[0].shape[2]
        else:
            max_len = cache_kvs[0].shape[3]
        if S + max_length > max_len:
            raise ValueError(f"cache max_len {max_len} < prompt {S} + max_length {max_length}")
        # the rotary tables must cover every position the decode steps will index (they have max_position_embeddings rows at
        # construction; the KV cache may be longer): grow them now, before any kernel or graph captures their pointers
        tb = getattr(self, "transformer_block", None)
        if tb is not None and hasattr(tb, "ensure_rope"):
            tb.ensure_rope(S + max_length)
        eos = torch.tensor([eos_token_id] if isinstance(eos_token_id, int) else list(eos_token_id or [-1]),
                           dtype=torch.int64, device=dev)
        st = dict(
            stop_flags=torch.zeros(B, dtype=torch.bool, device=dev), step_idx=torch.zeros(B, dtype=torch.int64, device=dev),
            max_dec_len=torch.full((B,), max_length, dtype=torch.int64, device=dev),
            min_dec_len=torch.full((B,), min_length, dtype=torch.int64, device=dev),
            seq_len_decoder=enc.clone(), pre_ids=torch.full((B, max_length + 1), -1, dtype=torch.int64, device=dev),
            eos=eos, out=torch.full((B, max_length), -1, dtype=torch.int64, device=dev),
            stop_count=torch.zeros(1, dtype=torch.int32, device=dev), col=torch.zeros(1, dtype=torch.int64, device=dev),
            penalty=torch.full((B,), penalty_score, dtype=torch.float32, device=dev),
            frequency=torch.full((B,), frequency_score, dtype=torch.float32, device=dev),
            presence=torch.full((B,), presence_score, dtype=torch.float32, device=dev),
            temperature=torch.full((B,), temperature, dtype=torch.float32, device=dev),
            plain=(penalty_score == 1.0 and frequency_score == 0.0 and presence_score == 0.0 and min_length <= 0
                   and temperature == 1.0),
            top_p=None, generator=None,
        )
        if top_p:
            st["top_p"] = torch.full((B,), float(top_p), dtype=torch.float32, device=dev)
            if seed:
                st["generator"] = torch.Generator(device=dev)
                st["generator"].manual_seed(int(seed))

        def update(next_tokens):
            # step_idx / stop flags / pre_ids / seq_len_decoder / token log — one kernel, no host sync.
            # seq_len_decoder is advanced for running sequences AFTER their token was chosen, so that the next
            # decode step appends at the right cache position.
            ops.generate_step_update(next_tokens, st["stop_flags"], st["step_idx"], st["max_dec_len"], st["seq_len_decoder"],
                                     st["pre_ids"], st["eos"], st["out"], st["stop_count"], out_col_dev=st["col"])
            if token_stream is not None:
                token_stream.push(next_tokens, st["stop_count"])

        if token_stream is not None:
            token_stream.reset(total_steps=max_length)
        # ---- prefill ("encoder" step) ----
        logits = self._prefill(ids, enc, cache_kvs)              # [B, V], last valid position of each prompt
        tgt = self._choose(logits, st)
        # the prompt already occupies enc[b] cache slots; the first generated token will be appended at slot enc[b]
        st["seq_len_decoder"] -= 1                                # update() adds 1 for running sequences
        update(tgt)

        # ---- decode loop ----
        def step():
            lg = self._decode(tgt, st["seq_len_decoder"], cache_kvs)
            nxt = self._choose(lg, st)
            tgt.copy_(nxt)
            update(tgt)

        from ... import _lib

        graph = None
        n_steps = max_length - 1
        done = 0
        # programmatic dependent launch: every GEMM of the decode step prefetches its weight tiles while the previous
        # kernel drains (b200_set_pdl); restored on exit
        old_pdl = _lib.load().b200_set_pdl(1 if use_pdl else 0)
        try:
            if use_cuda_graph and n_steps > 2:
                step(); done += 1                                # warm-up (sets kernel attributes, allocator pools)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                if st["generator"] is not None:                  # philox offsets of a private generator advance per replay
                    graph.register_generator_state(st["generator"])
                with torch.cuda.graph(graph):                    # capture only records: no step is executed here
                    step()
            while done < n_steps:
                if graph is not None:
                    graph.replay()
                else:
                    step()
                done += 1
                if sync_interval > 0 and done % sync_interval == 0 and int(st["stop_count"].item()) >= B:
                    break
        finally:
            _lib.load().b200_set_pdl(old_pdl)
        self.last_generate_steps = done + 1
        return st["out"], st["stop_flags"].to(torch.int32), st["seq_len_decoder"]
