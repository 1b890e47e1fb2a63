"""paddlenlp/utils/batch_sampler.py DistributedBatchSampler: batches of indices for one data-parallel rank.

Same constructor and iteration contract as the reference class (dataset, batch_size, num_replicas, rank, shuffle,
drop_last, consumed_samples; `set_epoch`): the sample order (optionally a per-epoch seeded permutation) is padded to a multiple
of num_replicas * batch_size by wrapping around, cut into global batches, and rank r takes slice r of every global batch."""
from __future__ import annotations

import math

import numpy as np

from .. import distributed as dist_env


class DistributedBatchSampler:
    def __init__(self, dataset, batch_size, num_replicas=None, rank=None, shuffle=False, drop_last=False, consumed_samples=0):
        assert isinstance(batch_size, int) and batch_size > 0, "batch_size should be a positive integer"
        self.dataset = dataset
        self.batch_size = batch_size
        self.nranks = int(num_replicas) if num_replicas is not None else dist_env.get_world_size()
        self.local_rank = int(rank) if rank is not None else dist_env.get_rank()
        self.shuffle = bool(shuffle)
        self.drop_last = bool(drop_last)
        self.epoch = 0
        self.consumed_samples = int(consumed_samples)
        self.num_samples = int(math.ceil(len(dataset) / self.nranks))
        self.total_size = self.num_samples * self.nranks

    def set_epoch(self, epoch: int = 0, consumed_samples: int = 0):
        self.epoch = int(epoch)
        self.consumed_samples = int(consumed_samples)

    def __iter__(self):
        n = len(self.dataset)
        indices = np.arange(n).tolist()
        if self.shuffle:
            np.random.RandomState(self.epoch).shuffle(indices)
            self.epoch += 1
        indices += indices[: self.total_size - len(indices)]
        indices = indices[self.consumed_samples:]
        glob = self.batch_size * self.nranks
        full = len(indices) // glob
        for g in range(full):
            chunk = indices[g * glob:(g + 1) * glob]
            yield chunk[self.local_rank * self.batch_size:(self.local_rank + 1) * self.batch_size]
        rest = indices[full * glob:]
        if rest and not self.drop_last:
            per = len(rest) // self.nranks
            if per:
                yield rest[self.local_rank * per:(self.local_rank + 1) * per]

    def __len__(self):
        n = self.num_samples - self.consumed_samples // self.nranks
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size
