"""Dataset helpers on the SFT path (paddlenlp/datasets): zero-padding (sample packing) for FlashMask attention."""
from .zero_padding_dataset import ZeroPadding, ZeroPaddingMapDataset, generate_greedy_packs

__all__ = ["ZeroPadding", "ZeroPaddingMapDataset", "generate_greedy_packs"]
