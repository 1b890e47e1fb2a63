"""DataCollatorForSeq2Seq — the right-padding collator of the SFT path (paddlenlp/data/data_collator.py:346-452 with
tokenizer.pad from transformers/tokenizer_utils_base.py:3230-3330): input_ids padded with pad_token_id, labels with
label_pad_token_id (-100), position_ids with 0, FlashMask `attn_mask_startend_row_indices` with 0 (:3256-3264)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


class DataCollatorForSeq2Seq:
    def __init__(self, tokenizer=None, model=None, padding="max_length", max_length: Optional[int] = None,
                 pad_to_multiple_of: Optional[int] = None, label_pad_token_id: int = -100, pad_token_id: Optional[int] = None,
                 return_tensors: str = "pt", return_attention_mask: Optional[bool] = None):
        self.tokenizer = tokenizer
        self.padding = padding
        self.max_length = max_length
        self.pad_to_multiple_of = pad_to_multiple_of
        self.label_pad_token_id = label_pad_token_id
        self.pad_token_id = pad_token_id if pad_token_id is not None else getattr(tokenizer, "pad_token_id", 0) or 0

    def __call__(self, features: List[Dict]) -> Dict[str, torch.Tensor]:
        longest = max(len(f["input_ids"]) for f in features)
        if self.padding == "max_length" and self.max_length is not None:
            if longest > self.max_length:
                raise ValueError(f"a sample of {longest} tokens exceeds max_length {self.max_length}")
            target = self.max_length
        else:
            target = longest
        if self.pad_to_multiple_of:
            target = (target + self.pad_to_multiple_of - 1) // self.pad_to_multiple_of * self.pad_to_multiple_of
        pads = {"input_ids": self.pad_token_id, "labels": self.label_pad_token_id, "position_ids": 0,
                "attn_mask_startend_row_indices": 0}
        out = {}
        for key, pad in pads.items():
            if key not in features[0]:
                continue
            rows = [list(f[key]) + [pad] * (target - len(f[key])) for f in features]
            out[key] = torch.tensor(rows, dtype=torch.int32 if key == "attn_mask_startend_row_indices" else torch.int64)
        return out
