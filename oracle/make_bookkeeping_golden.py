"""Extract the known-answer vectors of the reference's generation bookkeeping-op tests into tests/golden/bookkeeping.json.

    python oracle/make_bookkeeping_golden.py       (needs /root/reference; run in the build container)

Sources (PUBLIC test data of the reference; semantics are device independent):
    csrc/xpu/test/python/test_get_padding_offset_v2.py:24-66          csrc/xpu/test/python/test_update_inputs.py:22-95
    csrc/xpu/test/python/test_get_token_penalty_multi_scores_v2.py    csrc/xpu/test/python/test_set_stop_value_multi_ends_v2.py
    csrc/xpu/test/python/test_set_value_by_flags_and_idx_v2.py
Inputs drawn from numpy RNGs are regenerated with the seeds the tests use; expected arrays are parsed from the literals.
"""
import ast
import json
import os
import re

import numpy as np

REF = "/root/reference/csrc/xpu/test/python"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bookkeeping.json")


def literal_after(src, name, start=0):
    """First `name = <np.array|paddle.to_tensor>(<list literal>...` after `start`: returns (python list, end offset)."""
    m = re.compile(r"\b" + re.escape(name) + r"\s*=\s*(?:np\.array|paddle\.to_tensor)\(\s*").search(src, start)
    i = m.end()
    depth, j = 0, i
    while True:
        ch = src[j]
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return ast.literal_eval(src[i:j + 1]), j


def main():
    g = {}
    # ---- get_padding_offset_v2 ----
    src = open(f"{REF}/test_get_padding_offset_v2.py").read()
    np.random.seed(2023)
    max_len = 10
    seq_lens = np.array([4, 3, 6], "int32").reshape(-1, 1)
    cum_offset = np.cumsum((max_len - seq_lens).flatten(), -1, "int32")
    ids = np.zeros([3, max_len], "int64")
    for i in range(3):
        ids[i, 0:seq_lens[i, 0]] = np.random.randint(1, 10, seq_lens[i, 0], "int64")
    exp = {k: literal_after(src, k)[0] for k in ("ref_x_remove_padding", "ref_cum_offsets_out", "ref_padding_offset",
                                                 "ref_cu_seqlens_q", "ref_cu_seqlens_k")}
    g["get_padding_offset_v2"] = dict(input_ids=ids.tolist(), cum_offsets=cum_offset.tolist(), token_num=int(seq_lens.sum()),
                                      seq_lens=seq_lens.flatten().tolist(), **exp)
    # ---- token penalty v2 (both cases) ----
    src = open(f"{REF}/test_get_token_penalty_multi_scores_v2.py").read()
    # only the first case is asserted by the reference test (the second case's assert is commented out, :249-251)
    cases, pos = [], 0
    for _ in range(1):
        pre_ids, pos = literal_after(src, "pre_ids", pos)
        logits, pos = literal_after(src, "logits", pos)
        ref, pos = literal_after(src, "ref_logits", pos)
        cases.append(dict(pre_ids=pre_ids, logits=logits, ref_logits=ref))
    g["token_penalty_v2"] = dict(cases=cases, penalty_scores=[1.0, 1.0], frequency_scores=[0.1, 0.1], presence_scores=[0.0, 0.0],
                                 temperatures=[0.5, 0.25], bad_tokens=[0, 1], cur_len=[7, 6], min_len=[1, 8], eos_token_id=[2, 9])
    # ---- set_stop_value_multi_ends_v2 ----
    src = open(f"{REF}/test_set_stop_value_multi_ends_v2.py").read()
    np.random.seed(1)
    bs = 64
    stop_flags = np.random.randint(0, 2, [bs]).astype(bool)
    seq_lens = np.random.randint(0, 5, [bs]).astype("int32")
    g["set_stop_value_multi_ends_v2"] = dict(
        topk_ids=list(range(bs)), next_tokens=[0] * bs, stop_flags=stop_flags.tolist(), seq_lens=seq_lens.tolist(),
        end_ids=[0, 1, 2, 3, 4, 5], ref_topk_ids=literal_after(src, "ref_topk_ids")[0],
        ref_next_tokens=literal_after(src, "ref_next_tokens")[0], ref_stop_flags=literal_after(src, "ref_stop_flags")[0])
    # ---- set_value_by_flags_and_idx_v2 ----
    src = open(f"{REF}/test_set_value_by_flags_and_idx_v2.py").read()
    g["set_value_by_flags_and_idx_v2"] = dict(
        pre_ids_all=literal_after(src, "pre_ids_all")[0], input_ids=literal_after(src, "input_ids")[0], seq_lens_encoder=[1, 1],
        seq_lens_decoder=[1, 1], step_idx=[1, 1], stop_flags=[False, True], ref_pre_ids_all=literal_after(src, "ref_pre_ids_all")[0])
    # ---- update_inputs ----
    src = open(f"{REF}/test_update_inputs.py").read()
    np.random.seed(2023)
    bs, max_bs, max_input_length = 48, 64, 6144
    stop_flags = np.random.randint(0, 2, max_bs).astype("bool")
    this_time = np.zeros([bs], "int32"); enc = np.zeros([max_bs], "int32"); dec = np.zeros([max_bs], "int32")
    for i in range(bs):
        if i % 2 == 0:
            enc[i] = i; this_time[i] = i
        else:
            dec[i] = i; this_time[i] = 1
    input_ids = np.random.randint(1, 10, [max_bs, max_input_length], "int64")
    next_tokens = np.random.randint(1, 10, [max_bs], "int64")
    is_block_step = np.random.randint(0, 2, [max_bs]).astype("bool")
    g["update_inputs"] = dict(
        stop_flags=stop_flags.tolist(), seq_lens_this_time=this_time.tolist(), seq_lens_encoder=enc.tolist(),
        seq_lens_decoder=dec.tolist(), input_ids_col0_before=input_ids[:, 0].tolist(), input_ids_width=max_input_length,
        stop_nums=[max_bs], next_tokens=next_tokens.tolist(), is_block_step=is_block_step.tolist(),
        ref_not_need_stop=True, ref_seq_lens_this_time=literal_after(src, "ref_seq_lens_this_time_out")[0],
        ref_seq_lens_encoder=literal_after(src, "ref_seq_lens_encoder_out")[0],
        ref_seq_lens_decoder=literal_after(src, "ref_seq_lens_decoder_out")[0],
        ref_input_ids_col0=ast.literal_eval(re.search(r"input_ids_np\[:, 0\] = np\.array\((\[.*?\])", src, re.S).group(1)))
    with open(OUT, "w") as f:
        json.dump(g, f)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
