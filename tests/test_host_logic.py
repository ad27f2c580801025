"""CPU: host-side logic that needs no GPU - the live-key-block flags of the masked-block skipping (atlas_b200/ops.py), the chunk
tables of the multi-tensor optimiser kernels (atlas_b200/optim.py)."""
import numpy as np
import torch


def test_key_block_live_flags():
    from atlas_b200 import ops

    L = 200                                                   # ragged: 4 blocks of 64, the last one 8 keys wide
    lens = torch.tensor([200, 3, 64, 65, 130])
    mask = (torch.arange(L)[None, :] >= lens[:, None]).float() * -10000.0
    live = ops.key_block_live(mask)
    assert live.dtype == torch.uint8 and live.shape == (5, 4)
    assert live.tolist() == [[1, 1, 1, 1], [1, 0, 0, 0], [1, 0, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0]]
    # a hole of masked keys spanning one whole block; -1e9 masks (invert_attention_mask) count as dead as well
    hole = torch.zeros(1, 256)
    hole[0, 64:128] = -1e9
    assert ops.key_block_live(hole).tolist() == [[1, 0, 1, 1]]
    # small negative biases are live (only <= -5000 is treated as masked)
    assert bool(ops.key_block_live(torch.full((2, 128), -100.0)).all())
    # a fully masked row keeps every block (the reference's uniform-over-masked-keys softmax is computed, not skipped)
    assert bool(ops.key_block_live(torch.full((1, 192), -10000.0)).all())
    assert ops.key_block_live(None) is None


def test_optimizer_chunk_table_covers_every_element_once():
    from atlas_b200 import optim

    numels = [0, 1, optim.CHUNK - 1, optim.CHUNK, optim.CHUNK + 1, 5 * optim.CHUNK + 17]
    table = optim._chunk_table(numels)
    assert table.dtype == np.int32 and table.shape[1] == 2
    covered = {i: 0 for i in range(len(numels))}
    for t, c in table.tolist():
        begin, end = c * optim.CHUNK, min(numels[t], (c + 1) * optim.CHUNK)
        assert begin < end
        covered[t] += end - begin
    assert [covered[i] for i in range(len(numels))] == numels
    assert optim._ADAM_DESC.itemsize == 64 and optim._GRAD_DESC.itemsize == 24      # the C structs of include/atlas_b200.h
