"""paddlenlp.trainer surface kept by this build: Trainer, TrainingArguments, PdArgumentParser, TrainOutput."""
from .argparser import PdArgumentParser
from .trainer import PrinterCallback, TrainOutput, Trainer, TrainerCallback
from .training_args import TrainingArguments
