"""Weights interchange with the reference's checkpoints (convert.py:36-39, train.py:847-1054).

equinox serialises a model with `eqx.tree_serialise_leaves`: the array leaves of the pytree, in flattening order, each
written with `numpy.save` back to back into one file (non-array leaves -- the static hyper-parameter fields of Actor /
Critic -- are written too, as 0-d arrays).  The flattening order of `Model(actor, critic)` follows the field order of the
modules [R]: Actor: input_proj (weight, bias), rnns (per LSTMCell: weight_ih, weight_hh, bias), output_proj (weight, bias),
then the scalar fields; Critic likewise (train.py:847-1004).  [U]: xax wraps this stream in its own checkpoint archive
(`task.load_ckpt(..., part="model")`); this module reads / writes the leaf stream itself, which is what is inside.

No jax / equinox needed: numpy only.  `leaves_to_weights` returns the eqx-layout dicts `KbotStep.pack_weights` takes."""
from __future__ import annotations

import io
from pathlib import Path

import numpy as np


def read_leaves(path_or_bytes) -> list[np.ndarray]:
    """All `numpy.save` records of an eqx leaf stream, in order."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else Path(path_or_bytes).read_bytes()
    f = io.BytesIO(data)
    out = []
    while f.tell() < len(data):
        out.append(np.load(f, allow_pickle=False))
    return out


def write_leaves(leaves, path=None) -> bytes:
    f = io.BytesIO()
    for a in leaves:
        np.save(f, np.asarray(a), allow_pickle=False)
    b = f.getvalue()
    if path is not None:
        Path(path).write_bytes(b)
    return b


def _take_net(arrays: list[np.ndarray], pos: int, num_in: int, num_out: int, hidden: int, depth: int):
    """Consume one Actor / Critic from the array leaves (skipping 0-d static fields); shape asserts as train.py writes them."""
    def nxt():
        nonlocal pos
        while arrays[pos].ndim == 0:
            pos += 1
        a = arrays[pos]
        pos += 1
        return np.asarray(a, np.float32)

    w = {"w_in": nxt(), "b_in": nxt(), "layers": []}
    assert w["w_in"].shape == (hidden, num_in) and w["b_in"].shape == (hidden,), (w["w_in"].shape, w["b_in"].shape)
    for _ in range(depth):
        lw = {"w_ih": nxt(), "w_hh": nxt(), "b": nxt()}
        assert lw["w_ih"].shape == (4 * hidden, hidden) and lw["w_hh"].shape == (4 * hidden, hidden) and lw["b"].shape == (4 * hidden,)
        w["layers"].append(lw)
    w["w_out"], w["b_out"] = nxt(), nxt()
    assert w["w_out"].shape == (num_out, hidden) and w["b_out"].shape == (num_out,), (w["w_out"].shape, w["b_out"].shape)
    return w, pos


def leaves_to_weights(leaves: list[np.ndarray], hidden: int = 256, depth: int = 2) -> tuple[dict, dict]:
    """(actor, critic) eqx-layout weight dicts from the leaf stream of `Model` (train.py:1007-1054: actor first)."""
    actor, pos = _take_net(leaves, 0, 65, 40, hidden, depth)
    critic, pos = _take_net(leaves, pos, 475, 1, hidden, depth)
    assert all(a.ndim == 0 for a in leaves[pos:]), "unexpected array leaves after the critic"
    return actor, critic


def weights_to_leaves(actor: dict, critic: dict, static_actor=(), static_critic=()) -> list[np.ndarray]:
    """The inverse: array leaves in Model order (static scalar fields appended after each network when given)."""
    def net(w, static):
        out = [w["w_in"], w["b_in"]]
        for lw in w["layers"]:
            out += [lw["w_ih"], lw["w_hh"], lw["b"]]
        return out + [w["w_out"], w["b_out"]] + [np.asarray(s) for s in static]

    return net(actor, static_actor) + net(critic, static_critic)


def load_policy(path, hidden: int = 256, depth: int = 2) -> tuple[dict, dict]:
    """Checkpoint leaf stream -> (actor, critic) weights; feed them to `KbotStep.pack_weights` / `task.get_model`, then
    `KbotStep.policy_step` is convert.py's `step_fn` and a zero `[n, depth*2*hidden + 20]` array its `init_fn()` carry."""
    return leaves_to_weights(read_leaves(path), hidden, depth)
