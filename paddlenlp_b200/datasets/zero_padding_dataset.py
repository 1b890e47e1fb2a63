"""Zero padding = packing several tokenised samples into one max_length row (paddlenlp/datasets/zero_padding_dataset.py).

Behaviour restated from the reference:
* records longer than max_length are dropped (:127-128); in-order packing closes a pack when the next record would overflow
  it (:129-146); `greedy_zero_padding` buffers 500 records and places each, in order, into the pack with the most room (:18-39,147-170);
* a pack concatenates input_ids / labels, concatenates the per-sample position_ids, and shifts every sample's
  `attn_mask_startend_row_indices` by the tokens that precede it (:61-100) — with the SFT converters emitting
  `[seq_length] * seq_length` per sample (llm/utils/data.py:200-204) each key column's entry becomes the END of its sample;
* only the FlashMask form is produced here: the dense block-diagonal `attention_mask` alternative (:91-93,101-104) needs an
  S x S mask per row, which the flash path never materialises.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch


def generate_greedy_packs(examples: Sequence[Dict], max_length: int) -> List[List[Dict]]:
    """zero_padding_dataset.py:18-39: records in order; each goes to the open pack with the most room left (first such pack on
    ties), and a new pack is opened when it fits nowhere."""
    n = len(examples)
    if n == 0:
        return []
    left_len = [-1] * n
    left_len[0] = max_length
    packs: List[List[Dict]] = [[] for _ in range(n)]
    index, left_index = 0, 0
    while index < n:
        rec = examples[index]
        best = max(range(n), key=lambda i: (left_len[i], -i))        # np.argmax: first index of the maximum
        if len(rec["input_ids"]) <= left_len[best]:
            packs[best].append(rec)
            left_len[best] -= len(rec["input_ids"])
            index += 1
        else:
            left_index += 1
            left_len[left_index] = max_length
    return [p for p in packs if p]


class ZeroPadding:
    supported_input_keys = ["input_ids", "labels", "position_ids", "attn_mask_startend_row_indices"]

    @classmethod
    def _pad_batch_records(cls, batch_records: Sequence[Dict]) -> Dict[str, List[int]]:
        out: Dict[str, List[int]] = {"input_ids": [], "labels": [], "position_ids": [], "attn_mask_startend_row_indices": []}
        sequence_sum = 0
        for rec in batch_records:
            n = len(rec["input_ids"])
            if "labels" not in rec:
                raise ValueError("labels is required for ZeroPadding Dataset")
            out["input_ids"].extend(rec["input_ids"])
            out["labels"].extend(rec["labels"])
            out["position_ids"].extend(rec.get("position_ids", range(n)))
            rows = rec.get("attn_mask_startend_row_indices", [n] * n)
            out["attn_mask_startend_row_indices"].extend(int(i) + sequence_sum for i in rows)
            sequence_sum += n
        return out


class ZeroPaddingMapDataset(ZeroPadding, torch.utils.data.Dataset):
    def __init__(self, data, tokenizer=None, max_length: int = 2048, greedy_zero_padding: bool = False):
        self.tokenizer = tokenizer
        self.max_length = max_length
        self.greedy_zero_padding = greedy_zero_padding
        self.new_data = self._create_zero_padding_data(data)

    def _create_zero_padding_data(self, data):
        total = []
        if not self.greedy_zero_padding:
            batch, cur = [], 0
            for i in range(len(data)):
                rec = data[i]
                n = len(rec["input_ids"])
                if n > self.max_length:
                    continue
                if cur + n <= self.max_length:
                    batch.append(rec)
                    cur += n
                else:
                    total.append(self._pad_batch_records(batch))
                    batch, cur = [rec], n
            if batch:
                total.append(self._pad_batch_records(batch))
            return total
        buf: List[Dict] = []
        for i in range(len(data)):
            rec = data[i]
            if len(rec["input_ids"]) > self.max_length:
                continue
            if len(buf) < 500:
                buf.append(rec)
            else:
                total.extend(self._pad_batch_records(p) for p in generate_greedy_packs(buf, self.max_length))
                buf = [rec]
        if buf:
            total.extend(self._pad_batch_records(p) for p in generate_greedy_packs(buf, self.max_length))
        return total

    def __getitem__(self, idx):
        return self.new_data[idx]

    def __len__(self):
        return len(self.new_data)
