"""Streaming token output of the generation loop — the `save_output` / `get_output` / `read_res` trio of the reference
(csrc/gpu/save_with_output_msg.cc:28-52, csrc/gpu/get_output.cc:28-60, paddlenlp/utils/llm_utils.py:753-776) without the
per-step host synchronisation.

Reference: every decode step copies `next_tokens` and `not_need_stop` to the host (blocking) and `msgsnd`s
`{flag, bsz, tokens...}` into a SysV queue; a reader process loops `get_output(tensor, 0, wait_flag)` until flag == -1.
Here the decode step's stream writes that message into a ring of slots in pinned, device-mapped host memory
(b200_save_output_stream) and a reader polls the slot headers.  `get_output()` keeps the reference's contract:
returns `[flag, bsz, tok_0 .. tok_{bsz-1}]`, or `[-2, 0]` when nothing new has arrived (wait_flag=False)."""
from __future__ import annotations

import threading
import time
from typing import Callable, List, Optional

import torch

from ... import _lib
from ..._lib import ptr, stream_ptr

MAX_BSZ = 512          # llm/predict/predictor.py:49 — must match the message layout of save_output / get_output
HEADER = 3             # seq, flag, bsz


class TokenStreamOverrun(RuntimeError):
    pass


class TokenStream:
    def __init__(self, max_bsz: int = MAX_BSZ, num_slots: int = 4096, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("TokenStream needs a CUDA device (the producer is a kernel writing to mapped host memory)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.max_bsz, self.num_slots = int(max_bsz), int(num_slots)
        self.stride = HEADER + self.max_bsz
        self.ring = torch.zeros(self.num_slots, self.stride, dtype=torch.int32).pin_memory()   # host memory, device-visible (UVA)
        self._np = self.ring.numpy()                      # same memory: the reader never goes through CUDA
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.last_step = -1
        self._next = 0                                    # next step index the reader expects

    # ---- producer side (called from the generation loop; enqueues one kernel, never synchronises) ----
    def reset(self, total_steps: int = -1):
        """Start a new generation: forget old messages; `total_steps` (if known) marks the final message finished (-1)."""
        torch.cuda.current_stream(self.device).synchronize()
        self.ring.zero_()
        self.step_counter.zero_()
        self.last_step = int(total_steps) - 1 if total_steps and total_steps > 0 else -1
        self._next = 0

    def push(self, next_tokens: torch.Tensor, stop_count: Optional[torch.Tensor] = None):
        bs = next_tokens.numel()
        if bs > self.max_bsz:
            raise ValueError(f"batch {bs} > max_bsz {self.max_bsz}")
        _lib.call("b200_save_output_stream", ptr(next_tokens), ptr(stop_count), ptr(self.ring), self.stride, self.num_slots,
                  ptr(self.step_counter), self.last_step, bs, stream_ptr())

    # ---- consumer side (host only) ----
    def get_output(self, wait_flag: bool = False, timeout: float = 60.0) -> List[int]:
        """get_output(x, rank_id, wait_flag) of the reference: the next unread message, or [-2, 0] if none (non-blocking)."""
        slot = self._np[self._next % self.num_slots]
        want = self._next + 1
        t0 = time.time()
        while True:
            seq = int(slot[0])
            if seq == want:
                bsz = int(slot[2])
                msg = [int(slot[1]), bsz] + slot[HEADER:HEADER + bsz].tolist()
                if int(slot[0]) != want:                  # overwritten while reading: the reader fell a whole ring behind
                    raise TokenStreamOverrun(f"step {self._next} was overwritten before it was read")
                self._next += 1
                return msg
            if seq > want:
                raise TokenStreamOverrun(f"reader is at step {self._next} but slot already holds step {seq - 1}")
            if not wait_flag:
                return [-2, 0]
            if time.time() - t0 > timeout:
                raise TimeoutError(f"no token message for step {self._next} within {timeout} s")
            time.sleep(0.0002)

    def read_until_finished(self, on_message: Optional[Callable[[int, List[int]], None]] = None, timeout: float = 60.0):
        """The `read_res` loop (llm_utils.py:753-776): collect messages until flag == -1; returns [steps][bsz] token lists."""
        outputs = []
        while True:
            msg = self.get_output(True, timeout)
            outputs.append(msg[2:])
            if on_message is not None:
                on_message(len(outputs) - 1, msg)
            if msg[0] == -1:
                return outputs

    def start_reader(self, on_message: Optional[Callable[[int, List[int]], None]] = None, timeout: float = 60.0):
        """Run read_until_finished on a daemon thread; `.join()` the returned thread and read `.result` / `.error`."""
        th = threading.Thread(target=self._run, args=(on_message, timeout), daemon=True)
        th.result, th.error = None, None
        self._thread = th
        th.start()
        return th

    def _run(self, on_message, timeout):
        th = threading.current_thread()
        try:
            th.result = self.read_until_finished(on_message, timeout)
        except Exception as e:  # surfaced to the caller through .error
            th.error = e
