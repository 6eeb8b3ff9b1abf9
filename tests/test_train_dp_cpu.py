"""Data-parallel training host logic on CPU (BASELINE.json configs[4]): every rank owns its own
utterance piece, the only exchange is ONE summing all-reduce of the flat gradient per optimizer step, scaled by
1 / world.  World-size-2 gloo group on 127.0.0.1; the per-rank gradients come from the oracle (test infrastructure),
and the reduced vector must equal the gradient of the same two utterances processed as one batch by one process
(the loss terms are batch means: CRN_ELU.py:526-529, utility.py:222,913)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import SMALL
from oracle import crn_oracle, synth
from speech_enhancement_mi_b200 import training

L = 4000


def _oracle():
    w = synth.make_crn_weights(seed=7, **SMALL)
    return crn_oracle.CRNOracle({k: torch.from_numpy(v) for k, v in w.items()}, segment_length=3200, **SMALL)


def _flat(grads, names, layout, total):
    flat = torch.zeros(total)
    for n in names:
        if n in grads:
            flat[layout[n]:layout[n] + grads[n].numel()] = grads[n].reshape(-1)
    return flat


def _names_numels():
    shapes = synth.crn_param_shapes(**SMALL)
    names = [k for k in shapes if ".net.0." not in k]
    return names, [int(np.prod(shapes[k])) for k in names]


def test_flat_layout_is_contiguous():
    names, numels = _names_numels()
    layout, total = training.flat_layout(names, numels)
    assert total == sum(numels) and layout[names[0]] == 0
    for a, b, n in zip(names, names[1:], numels):
        assert layout[b] == layout[a] + n


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    names, numels = _names_numels()
    layout, total = training.flat_layout(names, numels)
    mix, src = synth.make_mixture(1, L, first_stream=rank)  # rank r trains on utterance r
    _, _, _, grads = crn_oracle.train_step_grads(_oracle(), mix, src, [L], False)
    flat = _flat(grads, names, layout, total)
    scale = training.allreduce_mean_(flat)
    assert scale == 1.0 / world
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), (flat * scale).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_allreduced_gradient_equals_single_process_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    names, numels = _names_numels()
    layout, total = training.flat_layout(names, numels)
    mix, src = synth.make_mixture(2, L)  # the same two utterances as one batch
    _, _, _, grads = crn_oracle.train_step_grads(_oracle(), mix, src, [L, L], False)
    want = _flat(grads, names, layout, total).numpy()
    assert np.abs(reduced - want).max() <= 2e-4 * np.abs(want).max()


def test_allreduce_without_process_group_is_identity():
    g = torch.arange(5.0)
    assert training.allreduce_mean_(g) == 1.0 and torch.equal(g, torch.arange(5.0))
