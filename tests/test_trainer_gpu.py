"""Trainer-level GPU tests: the reference's call pattern (Trainer(...).train() -> training_step -> loss.backward() ->
optimizer.step()) on the native engine, and one optimizer step checked against the oracle (rows a11, a13, a14)."""
import pytest
import torch

from oracle import llama_ref as R
from oracle import optim_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def tiny(model_type="llama"):
    import paddlenlp_b200.transformers as T

    kw = dict(vocab_size=512, hidden_size=256, intermediate_size=688, num_hidden_layers=2, num_attention_heads=2,
              num_key_value_heads=1, max_position_embeddings=256, seq_length=128, rope_theta=500000.0, rms_norm_eps=1e-5)
    return (T.Qwen2ForCausalLM(T.Qwen2Config(**kw)) if model_type == "qwen2" else T.LlamaForCausalLM(T.LlamaConfig(**kw)))


class ToyDataset(torch.utils.data.Dataset):
    def __init__(self, n, S, V, sft=False):
        g = torch.Generator().manual_seed(1234)
        self.tok = torch.randint(1, V, (n, S + 1), generator=g)
        self.sft = sft

    def __len__(self):
        return self.tok.shape[0]

    def __getitem__(self, i):
        ids, labels = self.tok[i, :-1].clone(), self.tok[i, 1:].clone()
        if self.sft:                                  # llm/utils/data.py:179-206: prompt tokens carry label -100
            labels[: 40 + i] = -100
        return {"input_ids": ids, "labels": labels}


@pytest.mark.parametrize("model_type,sft", [("llama", False), ("qwen2", True)])
def test_trainer_train_loop(model_type, sft, tmp_path):
    from paddlenlp_b200.trainer import Trainer, TrainingArguments

    model = tiny(model_type)
    args = TrainingArguments(output_dir=str(tmp_path), per_device_train_batch_size=2, gradient_accumulation_steps=2,
                             max_steps=8, learning_rate=2e-3, weight_decay=0.01, warmup_steps=1, logging_steps=2,
                             max_seq_length=128, lr_scheduler_type="linear")
    trainer = Trainer(model=model, args=args, train_dataset=ToyDataset(8, 128, 512, sft))
    out = trainer.train()
    hist = trainer.state.log_history
    assert out.global_step == 8 and len(hist) == 4
    assert hist[-1]["loss"] < hist[0]["loss"] - 0.3                     # 8 samples are memorised quickly
    for key in ("learning_rate", "global_step", "interval_samples_per_second", "interval_tokens_per_second_per_device",
                "interval_hardware_tflops_per_device"):
        assert key in hist[0]                                           # speed_metrics keys of trainer_utils.py:351-380


def test_one_optimizer_step_matches_oracle():
    """Gradient accumulation over 2 micro-batches + global-norm clip + AdamW vs the oracle restatement (a11-a13)."""
    from paddlenlp_b200.optimizer import AdamW, ClipGradByGlobalNorm

    cfg = R.RefConfig(vocab_size=512, hidden_size=256, intermediate_size=688, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=1, max_position_embeddings=256)
    w = R.init_weights(cfg, seed=21)
    w = {k: (v * 3).to(torch.bfloat16).float() if k.endswith("weight") and "norm" not in k else v for k, v in w.items()}
    model = tiny()
    model.set_state_dict(w)
    opt = AdamW(learning_rate=1e-3, weight_decay=0.1, grad_clip=ClipGradByGlobalNorm(1.0), engine=model.engine)
    g = torch.Generator().manual_seed(5)
    tok = torch.randint(0, 512, (4, 129), generator=g)
    grads_ref = None
    for mb in range(2):
        ids, labels = tok[2 * mb:2 * mb + 2, :-1].contiguous(), tok[2 * mb:2 * mb + 2, 1:].contiguous()
        loss, _ = model(input_ids=ids.to(DEV), labels=labels.to(DEV))
        (loss / 2).backward()
        _, _, gr = R.loss_and_grads(ids, labels, w, cfg, mode="bf16")
        grads_ref = {k: v / 2 for k, v in gr.items()} if grads_ref is None else {k: grads_ref[k] + gr[k] / 2 for k in gr}
    opt.step()
    # oracle: clip by the global norm of ALL gradients, then AdamW per tensor (decay only on matrices)
    flat = torch.cat([grads_ref[k].reshape(-1) for k in grads_ref])
    coef = float(optim_ref.clip_coef(flat, 1.0))
    new = model.engine.named_views(flat=opt.master)        # fp32 master weights (bf16 params = round(master))
    worst = 0.0
    for k, g_ in grads_ref.items():
        decay = torch.full_like(w[k], ("norm" not in k and "bias" not in k), dtype=torch.bool)
        p, _, _, _ = optim_ref.adamw_step(w[k], torch.zeros_like(w[k]), torch.zeros_like(w[k]), g_ * coef, lr=1e-3, beta1=0.9,
                                          beta2=0.999, eps=1e-8, weight_decay=0.1, step=1, decay_mask=decay, max_grad_norm=0.0)
        upd_ref, upd = p - w[k], new[k].detach().float().cpu() - w[k]
        # the first Adam step moves every weight with a gradient by ~lr * sign(g) (+ decay): a sign flip can only come
        # from a gradient whose sign differs, i.e. |g| below the bf16 gradient noise
        agree = (torch.sign(upd_ref) == torch.sign(upd)).float().mean().item()
        worst = max(worst, 1 - agree)
        assert agree > 0.95, (k, agree)
        rel = ((upd - upd_ref).norm() / (upd_ref.norm() + 1e-30)).item()
        assert rel < 0.35, (k, rel)
    gn = float(opt.grad_norm())
    assert abs(gn - float(flat.norm())) < 0.03 * float(flat.norm())


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """save at step 3 (model shards + optimizer.safetensors + master_weights.safetensors + scheduler +
    trainer_state.json + rng), resume in a fresh Trainer, finish at step 6 -> the same bits as the uninterrupted run."""
    import json
    import os

    from paddlenlp_b200.trainer import Trainer, TrainingArguments
    from paddlenlp_b200.trainer.trainer import get_last_checkpoint

    def run(out_dir, max_steps, resume=None, save_steps=0):
        model = tiny("qwen2")
        args = TrainingArguments(output_dir=str(out_dir), per_device_train_batch_size=2, gradient_accumulation_steps=2,
                                 max_steps=max_steps, learning_rate=1e-3, weight_decay=0.01, warmup_steps=2, logging_steps=1,
                                 max_seq_length=128, lr_scheduler_type="cosine", save_steps=save_steps, save_total_limit=1)
        tr = Trainer(model=model, args=args, train_dataset=ToyDataset(16, 128, 512, True))
        tr.train(resume_from_checkpoint=resume)
        return tr

    full = run(tmp_path / "full", 6)
    run(tmp_path / "part", 3, save_steps=1)                       # saves at 1, 2, 3; save_total_limit keeps only the last
    ck = get_last_checkpoint(str(tmp_path / "part"))
    assert ck.endswith("checkpoint-3") and sorted(os.listdir(tmp_path / "part")) == ["checkpoint-3"]
    files = set(os.listdir(ck))
    assert {"config.json", "model.safetensors", "optimizer.safetensors", "master_weights.safetensors", "scheduler.pdparams",
            "trainer_state.json", "rng_state.pth", "training_args.json"} <= files
    assert json.load(open(os.path.join(ck, "trainer_state.json")))["global_step"] == 3
    # resume: max_steps 6 with the same schedule; the data loader skips the 3*2 consumed batches
    import shutil
    shutil.copytree(ck, tmp_path / "resumed" / "checkpoint-3")
    res = run(tmp_path / "resumed", 6, resume=True)
    assert res.state.global_step == 6
    assert torch.equal(res.model.engine.flat_params, full.model.engine.flat_params)
    assert torch.equal(res.optimizer.master, full.optimizer.master)
    assert torch.equal(res.optimizer.exp_avg_sq, full.optimizer.exp_avg_sq)
    assert [h["loss"] for h in res.state.log_history][-3:] == [h["loss"] for h in full.state.log_history][-3:]
    assert res.optimizer.get_lr() == full.optimizer.get_lr()
