"""PdArgumentParser — dataclass-driven argument parsing (paddlenlp/trainer/argparser.py): command line and/or JSON."""
from __future__ import annotations

import argparse
import dataclasses
import json
import sys
from typing import Optional, Tuple, get_args, get_origin, Union


def _base_type(tp):
    if get_origin(tp) is Union:
        args = [a for a in get_args(tp) if a is not type(None)]
        return args[0] if args else str
    return tp


def _str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError(f"boolean expected, got {v}")


class PdArgumentParser(argparse.ArgumentParser):
    def __init__(self, dataclass_types, **kwargs):
        super().__init__(**kwargs)
        if dataclasses.is_dataclass(dataclass_types):
            dataclass_types = [dataclass_types]
        self.dataclass_types = list(dataclass_types)
        import typing

        seen = set()
        for dt in self.dataclass_types:
            hints = typing.get_type_hints(dt)
            for f in dataclasses.fields(dt):
                if f.name in seen:          # a name shared by two argument classes (e.g. max_seq_length: DataArguments and this
                    continue                # build's TrainingArguments) is parsed once and handed to both
                seen.add(f.name)
                tp = _base_type(hints.get(f.name, str))
                kw = {}
                if tp is bool:
                    kw.update(type=_str2bool, nargs="?", const=True)
                elif tp in (int, float, str):
                    kw.update(type=tp)
                default = f.default if f.default is not dataclasses.MISSING else None
                self.add_argument(f"--{f.name}", default=default, **kw)

    def _build(self, values: dict) -> Tuple:
        outs = []
        for dt in self.dataclass_types:
            keys = {f.name for f in dataclasses.fields(dt)}
            outs.append(dt(**{k: v for k, v in values.items() if k in keys}))
        return tuple(outs)

    def parse_args_into_dataclasses(self, args=None):
        ns = self.parse_args(args)
        return self._build(vars(ns))

    def parse_json_file(self, json_file: str):
        with open(json_file) as f:
            data = json.load(f)
        return self.parse_dict(data)

    def parse_dict(self, data: dict):
        ns = vars(self.parse_args([]))
        ns.update(data)
        return self._build(ns)

    def parse_json_file_and_cmd_lines(self):
        """`script.py config.json --override value` (llm/run_pretrain.py:359-365)."""
        argv = sys.argv[1:]
        if argv and argv[0].endswith(".json"):
            with open(argv[0]) as f:
                data = json.load(f)
            ns = vars(self.parse_args(argv[1:]))
            explicit = {a.lstrip("-").split("=")[0] for a in argv[1:] if a.startswith("--")}
            for k, v in data.items():
                if k not in explicit:
                    ns[k] = v
            return self._build(ns)
        return self.parse_args_into_dataclasses(argv)
