"""A small continuous-batching simulation that drives step_paddle (oracle or CUDA op) with a deliberately tight block pool, so
that every branch fires: freeing, on-demand allocation, pre-emption of the largest holder, recovery.  Shared by the CPU
invariant test of the oracle and the GPU bit-exactness test of the kernel."""
import numpy as np

KEYS_I32 = ("seq_lens_this_time", "ori_seq_lens_encoder", "seq_lens_encoder", "seq_lens_decoder", "block_tables",
            "encoder_block_lens", "step_block_list", "step_lens", "recover_block_list", "recover_lens", "need_block_list",
            "need_block_len", "used_list_len", "free_list", "free_list_len")
KEYS_BOOL = ("stop_flags", "is_block_step")
KEYS_I64 = ("input_ids", "pre_ids", "step_idx", "next_tokens")
ORDER = ("stop_flags", "seq_lens_this_time", "ori_seq_lens_encoder", "seq_lens_encoder", "seq_lens_decoder", "block_tables",
         "encoder_block_lens", "is_block_step", "step_block_list", "step_lens", "recover_block_list", "recover_lens",
         "need_block_list", "need_block_len", "used_list_len", "free_list", "free_list_len", "input_ids", "pre_ids", "step_idx",
         "next_tokens")


def make_state(seed=0, bsz=6, block_size=4, block_num_per_seq=10, length=40, num_blocks=22, max_dec=24):
    rng = np.random.RandomState(seed)
    prompt = rng.randint(3, 13, size=bsz).astype(np.int32)
    st = {
        "stop_flags": np.zeros(bsz, bool), "is_block_step": np.zeros(bsz, bool),
        "seq_lens_this_time": prompt.copy(), "ori_seq_lens_encoder": prompt.copy(), "seq_lens_encoder": prompt.copy(),
        "seq_lens_decoder": np.zeros(bsz, np.int32), "block_tables": np.full((bsz, block_num_per_seq), -1, np.int32),
        "encoder_block_lens": np.zeros(bsz, np.int32), "step_block_list": np.full(bsz, -1, np.int32), "step_lens": np.zeros(1, np.int32),
        "recover_block_list": np.full(bsz, -1, np.int32), "recover_lens": np.zeros(1, np.int32),
        "need_block_list": np.full(bsz, -1, np.int32), "need_block_len": np.zeros(1, np.int32),
        "used_list_len": np.zeros(bsz, np.int32), "free_list": np.full(num_blocks, -1, np.int32), "free_list_len": np.zeros(1, np.int32),
        "input_ids": rng.randint(5, 1000, size=(bsz, length)).astype(np.int64),
        "pre_ids": np.full((bsz, max_dec + 1), -1, np.int64), "step_idx": np.zeros(bsz, np.int64),
        "next_tokens": np.full(bsz, -1, np.int64),
    }
    free = list(range(num_blocks))
    for b in range(bsz):                               # prompt ("encoder") blocks come from the tail of the free list
        n = (int(prompt[b]) + block_size - 1) // block_size
        for j in range(n):
            st["block_tables"][b, j] = free.pop()
        st["encoder_block_lens"][b] = n
    st["free_list"][:len(free)] = free
    st["free_list_len"][0] = len(free)
    return st, rng


def between_steps(st, rng, block_size, max_dec, p_stop=0.06):
    """What the model + update_inputs do between two step_paddle calls, reduced to the fields step_paddle reads."""
    bsz = st["stop_flags"].shape[0]
    for b in range(bsz):
        if st["stop_flags"][b]:
            continue
        if st["seq_lens_encoder"][b] > 0:              # a (re-)prefill just ran: the whole sequence is now in the cache
            st["seq_lens_decoder"][b] = st["seq_lens_encoder"][b]
            st["seq_lens_encoder"][b] = 0
            if st["step_idx"][b] == 0:
                # first prefill of a fresh request: blocks beyond the prompt count as decoder blocks from now on
                pass
        else:
            st["seq_lens_decoder"][b] += 1
        tok = int(rng.randint(5, 1000))
        st["next_tokens"][b] = tok
        st["step_idx"][b] += 1
        if st["step_idx"][b] < st["pre_ids"].shape[1]:
            st["pre_ids"][b, st["step_idx"][b]] = tok
        st["seq_lens_this_time"][b] = 1
        if st["step_idx"][b] >= max_dec or rng.rand() < p_stop:
            st["stop_flags"][b] = True
            st["seq_lens_this_time"][b] = 0


def check_invariants(st, num_blocks):
    fl = st["free_list"][: int(st["free_list_len"][0])].tolist()
    held = [int(x) for x in st["block_tables"].reshape(-1) if x >= 0]
    allb = sorted(fl + held)
    assert allb == list(range(num_blocks)), ("every cache block is owned exactly once", allb)
    assert int(st["need_block_len"][0]) == 0 and int(st["recover_lens"][0]) == 0
    for b in range(st["stop_flags"].shape[0]):
        if not st["stop_flags"][b] and st["seq_lens_decoder"][b] != 0:
            # a running sequence always owns the block its next token lands in
            assert st["block_tables"][b, st["seq_lens_decoder"][b] // BLOCK_SIZE_FOR_CHECK[0]] >= 0, b


BLOCK_SIZE_FOR_CHECK = [4]
