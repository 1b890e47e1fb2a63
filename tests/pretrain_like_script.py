"""A pre-training script written the way llm/run_pretrain.py:14-47,358-575 is: the SAME `paddlenlp.*` import lines, the same
argument dataclasses, config / model / scheduler / Trainer construction and `trainer.train()` call pattern — run UNCHANGED
against this repo through the `paddlenlp` import shim (tests/test_trainer_gpu.py::test_reference_style_script_runs_unchanged).
The two things a hot-path build cannot provide are replaced inline and marked: the `paddle` import and the mmap'd
token-file dataset (`paddlenlp.data.causal_dataset`), which becomes a synthetic token dataset of the same item format."""
import math
import os
import sys
import time
from dataclasses import dataclass, field
from typing import Optional

import torch  # <- `import paddle` in the reference

from paddlenlp.trainer import (
    PdArgumentParser,
    Trainer,
    TrainingArguments,
    get_last_checkpoint,
    set_seed,
    speed_metrics,
)
from paddlenlp.transformers import (
    AutoConfig,
    AutoModelForCausalLM,
    AutoModelForCausalLMPipe,
    AutoTokenizer,
    CosineAnnealingWithWarmupDecay,
    LinearAnnealingWithWarmupDecay,
    register_sequence_parallel_allreduce_hooks,
)
from paddlenlp.transformers.configuration_utils import LlmMetaConfig, llmmetaclass
from paddlenlp.utils.batch_sampler import DistributedBatchSampler
from paddlenlp.utils.log import logger
from paddlenlp.utils.tools import get_env_device

os.environ["USE_CASUAL_MASK"] = "True"


@dataclass
@llmmetaclass
class PreTrainingArguments(TrainingArguments):
    min_learning_rate: float = field(default=1e-5, metadata={"help": "Minimum learning rate deacyed to."})
    decay_steps: float = field(default=None, metadata={"help": "The steps use to control the learing rate."})


@dataclass
class DataArguments:
    input_dir: str = field(default=None, metadata={"help": "The name of the dataset to use (via the datasets library)."})
    split: str = field(default="949,50,1", metadata={"help": "Train/valid/test data split."})
    max_seq_length: int = field(default=1024, metadata={"help": "The maximum total input sequence length after tokenization."})


@dataclass
class ModelArguments:
    model_name_or_path: str = field(default="__internal_testing__/tiny-random-llama")
    tokenizer_name_or_path: Optional[str] = field(default=None)
    num_hidden_layers: Optional[int] = field(default=None, metadata={"help": "num_hidden_layers."})
    continue_training: bool = field(default=False)
    fuse_attention_qkv: bool = field(default=None)
    fuse_attention_ffn: bool = field(default=None)


class SyntheticTokenDataset(torch.utils.data.Dataset):   # <- build_train_valid_test_datasets (mmap'd .bin/.idx files) in the reference
    def __init__(self, n, seq_len, vocab):
        self.tok = torch.randint(1, vocab, (n, seq_len + 1), generator=torch.Generator().manual_seed(1234))

    def __len__(self):
        return self.tok.shape[0]

    def __getitem__(self, i):
        # the reference's _collate_data (run_pretrain.py:245-255): input_ids = tokens[:-1], labels = tokens[1:]
        return {"input_ids": self.tok[i, :-1].clone(), "labels": self.tok[i, 1:].clone()}


class PretrainingTrainer(Trainer):
    def _get_train_sampler(self):                         # run_pretrain.py:341-349 — keeps the file order
        return DistributedBatchSampler(
            self.train_dataset,
            batch_size=self.args.per_device_train_batch_size,
            shuffle=False,
            num_replicas=self.args.dataset_world_size,
            rank=self.args.dataset_rank,
            drop_last=self.args.dataloader_drop_last,
        )


def main():
    parser = PdArgumentParser((ModelArguments, DataArguments, PreTrainingArguments))
    if len(sys.argv) >= 2 and sys.argv[1].endswith(".json"):
        model_args, data_args, training_args = parser.parse_json_file_and_cmd_lines()
    else:
        model_args, data_args, training_args = parser.parse_args_into_dataclasses()

    if training_args.no_recompute_layers is not None:
        training_args.no_recompute_layers.sort()

    if model_args.tokenizer_name_or_path is None:
        model_args.tokenizer_name_or_path = model_args.model_name_or_path

    set_seed(seed=training_args.seed)
    training_args.print_config(model_args, "Model")
    training_args.print_config(data_args, "Data")
    logger.warning(
        f"Process rank: {training_args.local_rank}, device: {training_args.device}, world_size: {training_args.world_size}, "
        + f"distributed training: {bool(training_args.local_rank != -1)}, 16-bits training: {training_args.fp16 or training_args.bf16}"
    )

    last_checkpoint = None
    if os.path.isdir(training_args.output_dir) and training_args.do_train and not training_args.overwrite_output_dir:
        last_checkpoint = get_last_checkpoint(training_args.output_dir)
        if last_checkpoint is not None and training_args.resume_from_checkpoint is None:
            logger.info(f"Checkpoint detected, resuming training at {last_checkpoint}.")

    config = AutoConfig.from_pretrained(model_args.model_name_or_path)
    LlmMetaConfig.set_llm_config(config, training_args)
    config.seq_length = data_args.max_seq_length
    if not model_args.continue_training:
        config.max_position_embeddings = max(config.max_position_embeddings, data_args.max_seq_length)
    config.num_hidden_layers = (
        model_args.num_hidden_layers if model_args.num_hidden_layers is not None else config.num_hidden_layers
    )
    if model_args.fuse_attention_qkv is not None:
        config.fuse_attention_qkv = model_args.fuse_attention_qkv
    if model_args.fuse_attention_ffn is not None:
        config.fuse_attention_ffn = model_args.fuse_attention_ffn
    assert config.num_attention_heads % config.sep_parallel_degree == 0
    assert config.seq_length % config.context_parallel_degree == 0
    print("Final pre-training config:", config)

    dtype = "float32"
    if training_args.fp16_opt_level == "O2":
        if training_args.fp16:
            dtype = "float16"
        if training_args.bf16:
            dtype = "bfloat16"

    model_class = AutoModelForCausalLM
    if training_args.pipeline_parallel_degree > 1:
        model_class = AutoModelForCausalLMPipe
    if model_args.continue_training:
        model = model_class.from_pretrained(model_args.model_name_or_path, config=config, dtype=dtype)
    else:
        model = model_class.from_config(config, dtype=dtype)

    if training_args.sequence_parallel:
        register_sequence_parallel_allreduce_hooks(
            model, training_args.gradient_accumulation_steps, training_args.fuse_sequence_parallel_allreduce
        )
    if training_args.recompute:
        model.recompute_enable()

    if training_args.decay_steps is None:
        training_args.decay_steps = training_args.max_steps
    if training_args.warmup_steps > 0:
        warmup_steps = training_args.warmup_steps
    else:
        warmup_steps = training_args.warmup_ratio * training_args.max_steps

    lr_scheduler = None
    if training_args.lr_scheduler_type.value == "cosine":
        lr_scheduler = CosineAnnealingWithWarmupDecay(
            max_lr=training_args.learning_rate,
            min_lr=training_args.min_learning_rate,
            warmup_step=warmup_steps,
            decay_step=training_args.decay_steps,
            last_epoch=0,
        )
    elif training_args.lr_scheduler_type.value == "linear":
        lr_scheduler = LinearAnnealingWithWarmupDecay(
            max_lr=training_args.learning_rate,
            min_lr=training_args.min_learning_rate,
            warmup_step=warmup_steps,
            decay_step=training_args.decay_steps,
            last_epoch=0,
        )

    # 8 samples seen repeatedly (several epochs within max_steps): random tokens can only be memorised, which is what makes the
    # loss fall in a few steps
    train_dataset = SyntheticTokenDataset(8 * training_args.dataset_world_size, data_args.max_seq_length, config.vocab_size)

    trainer = PretrainingTrainer(
        model=model,
        args=training_args,
        data_collator=None,
        train_dataset=train_dataset if training_args.do_train else None,
        eval_dataset=None,
        optimizers=(None, lr_scheduler),
        tokenizer=None,
    )

    checkpoint = None
    if training_args.resume_from_checkpoint is not None:
        checkpoint = training_args.resume_from_checkpoint
    elif last_checkpoint is not None:
        checkpoint = last_checkpoint

    if training_args.do_train:
        train_result = trainer.train(resume_from_checkpoint=checkpoint)
        metrics = train_result.metrics
        if not int(os.getenv("test_ci_no_save_model", 0)):
            trainer.save_model()
        trainer.log_metrics("train", metrics)
        trainer.save_metrics("train", metrics)
        trainer.save_state()
        print("FINAL_LOSS_HISTORY", [h["loss"] for h in trainer.state.log_history])


if __name__ == "__main__":
    main()
