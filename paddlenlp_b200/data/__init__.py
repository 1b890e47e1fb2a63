from .data_collator import DataCollatorForSeq2Seq

__all__ = ["DataCollatorForSeq2Seq"]
