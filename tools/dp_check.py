"""2+ GPU check of the data-parallel Trainer path (torchrun): overlapped vs single-call gradient exchange give the same
model, and every rank ends with identical parameters.

    python -m torch.distributed.run --nproc-per-node 2 tools/dp_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import paddlenlp_b200.transformers as T  # noqa: E402
from paddlenlp_b200 import distributed as D  # noqa: E402
from paddlenlp_b200.trainer import Trainer, TrainingArguments  # noqa: E402


class Toy(torch.utils.data.Dataset):
    def __init__(self, n, S, V):
        g = torch.Generator().manual_seed(7)
        self.tok = torch.randint(1, V, (n, S + 1), generator=g)

    def __len__(self):
        return self.tok.shape[0]

    def __getitem__(self, i):
        return {"input_ids": self.tok[i, :-1].clone(), "labels": self.tok[i, 1:].clone()}


def run(overlap: bool):
    torch.manual_seed(0)
    cfg = T.LlamaConfig(vocab_size=1024, hidden_size=512, intermediate_size=1376, num_hidden_layers=4, num_attention_heads=4,
                        num_key_value_heads=2, max_position_embeddings=256, seq_length=256)
    model = T.LlamaForCausalLM(cfg)
    args = TrainingArguments(output_dir="/tmp/dp_check", per_device_train_batch_size=2, gradient_accumulation_steps=3, max_steps=4,
                             learning_rate=1e-3, weight_decay=0.01, warmup_steps=1, logging_steps=1, max_seq_length=256,
                             lr_scheduler_type="linear")
    wrapped = D.DataParallel(model, overlap=overlap)
    tr = Trainer(model=wrapped, args=args, train_dataset=Toy(96, 256, 1024))
    out = tr.train()
    return model.engine.flat_params.clone(), [h["loss"] for h in tr.state.log_history], out


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    D.init_parallel_env("nccl")
    pa, la, _ = run(True)
    pb, lb, _ = run(False)
    lo, hi = pa.float().clone(), pa.float().clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same_across_ranks = bool(torch.equal(lo, hi))
    rel = ((pa.float() - pb.float()).norm() / pb.float().norm()).item()
    if D.get_rank() == 0:
        print(json.dumps({"world": D.get_world_size(), "params_identical_across_ranks": same_across_ranks,
                          "overlap_vs_single_call_rel_diff": rel, "loss_overlap": la, "loss_single": lb}), flush=True)
    assert same_across_ranks and rel < 2e-3 and max(abs(a - b) for a, b in zip(la, lb)) < 5e-3


if __name__ == "__main__":
    main()
