"""The fused MLP kernel (csrc/mlp_fused.cu) follows a static, compile-time plan of its 3-slot weight ring: which slot
every MMA step reads, whether it must wait for a fresh load first, and when a slot is released for the producer.
A wrong plan deadlocks or silently multiplies by stale weights, so the plan is replayed here on the host (no GPU):
producer and MMA issuer are simulated against the slot states for three consecutive tile pairs of every program."""
import ctypes

import pytest

from panonerf_b200 import _lib

PROGRAMS = {0: "forward", 1: "forward + Jacobian sweep", 2: "backward dgrad chain", 3: "adjoint sweep"}


def _plan(prog):
    buf = (ctypes.c_longlong * 8000)()
    assert _lib.lib().pnb_mlp_fused_plan(prog, buf, 8000) == 0
    n_loads, n_steps = buf[0], buf[1]
    loads = [tuple(buf[2 + 4 * i + j] for j in range(4)) for i in range(n_loads)]
    base = 2 + 4 * n_loads
    steps = [tuple(buf[base + 12 * i + j] for j in range(12)) for i in range(n_steps)]
    return loads, steps


@pytest.mark.parametrize("prog", sorted(PROGRAMS))
def test_ring_plan_is_consistent(prog):
    loads, steps = _plan(prog)
    assert loads and steps
    content, state = [None] * 3, ["empty"] * 3
    all_loads, all_steps = loads * 3, steps * 3          # the plan repeats verbatim for every tile pair
    li = 0
    for si, (t, sidx, off, aenc, ws, wl, wr, es, el, er, first, op) in enumerate(all_steps):
        # the producer runs ahead as far as free slots allow, strictly in plan order
        while li < len(all_loads) and state[all_loads[li][0]] == "empty":
            slot, is_enc, lt, loff = all_loads[li]
            content[slot] = ("enc", lt) if is_enc else ("w", loff)
            state[slot] = "loaded"
            li += 1
        if aenc:                                          # step multiplies the IPE tile of (op, tile t)
            want = "loaded" if el else "inuse"
            assert state[es] == want and content[es] == ("enc", t), (PROGRAMS[prog], si)
            state[es] = "inuse"
        want = "loaded" if wl else "inuse"
        assert state[ws] == want and content[ws] == ("w", off), (PROGRAMS[prog], si, content[ws], state[ws])
        state[ws] = "inuse"
        if wr:
            state[ws] = "empty"
        if aenc and er:
            state[es] = "empty"
    assert li == len(all_loads) and state == ["empty"] * 3      # every load consumed, every slot handed back


@pytest.mark.parametrize("prog", sorted(PROGRAMS))
def test_first_step_of_every_op_overwrites_the_accumulator(prog):
    _, steps = _plan(prog)
    seen = set()
    for (t, sidx, off, aenc, ws, wl, wr, es, el, er, first, op) in steps:
        assert bool(first) == ((op, t) not in seen), (PROGRAMS[prog], op, t)
        seen.add((op, t))


def test_weight_reuse_saves_loads():
    """Tile 1 walks the k-blocks of an op backwards, so fewer than one load per MMA step is needed."""
    for prog in PROGRAMS:
        loads, steps = _plan(prog)
        assert len(loads) < len(steps), PROGRAMS[prog]
