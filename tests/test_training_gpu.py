"""One miniature generation of train() on the GPU: self-play -> device replay buffer -> AdamW steps -> weights back."""
import numpy as np
import pytest
import torch

import alphazero_chess_b200 as az
from alphazero_chess_b200 import training as tr
from helpers import orc

pytestmark = pytest.mark.gpu


def test_generation_loop_runs_and_learns():
    torch.manual_seed(0)
    model = tr.import_weights(tr.AlphaZeroNet(), az.random_weights(seed=42)).cuda()
    opt = tr.make_optimizer(model)
    with az.Engine(max_games=32, num_simulations=16, seed=1) as eng:
        replay = az.ReplayBuffer(eng, capacity=20_000, max_batch=64)
        before, _ = (lambda e: (e.load_weights(tr.export_weights(model)), e.forward(orc.startpos()))[1])(eng)
        m0 = tr.run_generation(eng, replay, model, opt, iteration=0, n_games=32, min_replay_size=200, num_steps=6, batch_size=64)
        assert m0["trained"] and m0["positions"] > 300 and 0 < m0["new_unique_states"] <= m0["positions"]
        assert m0["replay_buffer_size"] == len(replay) and np.isfinite(m0["avg_policy_loss"]) and np.isfinite(m0["avg_value_loss"])
        after, _ = eng.forward(orc.startpos())
        assert not np.array_equal(before, after)          # the engine really runs on the updated weights
        # the engine (bf16 tcgen05 path) and the torch model (fp32, eval mode) agree on the new weights
        model.eval()
        with torch.no_grad():
            tp, tv = model(torch.from_numpy(orc.to_tensor(orc.startpos())[None]).cuda())
        assert np.abs(tp.cpu().numpy()[0] - after[0]).max() < 1e-2
        m1 = tr.run_generation(eng, replay, model, opt, iteration=1, n_games=32, min_replay_size=200, num_steps=6, batch_size=64)
        assert m1["trained"] and m1["replay_buffer_size"] >= m0["replay_buffer_size"]
        assert abs(m1["learning_rate"] - tr.get_cyclical_lr(1)) < 1e-12
        replay.close()
