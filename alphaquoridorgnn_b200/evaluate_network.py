"""New-parameter evaluation -- drop-in for reference evaluate_network.py:14-98: EN_GAME_COUNT games of
latest vs best with PV-MCTS at temperature 1.0, colours alternating, promotion if the average point
exceeds 0.5.  All games run in lock-step; at every ply the active games are split by which model is to
move and each half is searched with that model's batched evaluator."""
import os
from shutil import copy

import torch

from . import game_logic as gl
from . import pv_mcts
from .constants import PV_NETWORK_PATH
from .pv_network_gnn import GNNNetwork
from .positions import start_states

EN_GAME_COUNT = 15    # evaluate_network.py:14
EN_TEMPERATURE = 1.0  # evaluate_network.py:15


def first_player_point(ended_state):
    # 1: first player wins, 0: first player loses, 0.5: draw (evaluate_network.py:18-22)
    if ended_state.is_lose():
        return 0 if ended_state.is_first_player() else 1
    return 0.5


def update_best_player(model_dir=PV_NETWORK_PATH):
    copy(os.path.join(model_dir, 'latest.pth'), os.path.join(model_dir, 'best.pth'))
    print('Latest model is better than current best. Replacing best model with latest.')


@torch.no_grad()
def play_matches(model0, model1, num_games=EN_GAME_COUNT, temperature=EN_TEMPERATURE, sims=None, seed=0, device=None,
                 details=False):
    """-> total points of model0 over num_games (game i: model0 moves first iff i is even,
    evaluate_network.py:69-73).  details=True: -> (total, per-game first_player_point list, per-game action lists)."""
    dev = gl._dev(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    sims = sims or pv_mcts.PV_EVALUATE_COUNT
    searchers = [pv_mcts.searcher_for(m, sims, dev) for m in (model0, model1)]
    states = start_states(num_games, dev)
    gid = torch.arange(num_games, device=dev)
    points = torch.zeros(num_games, dtype=torch.float64, device=dev)  # points of the FIRST player of each game
    log = [[] for _ in range(num_games)] if details else None
    ply = 0
    while states.shape[0] > 0:
        # first player moves on even plies; game i's first player is model (i % 2)
        mover = (gid + ply) % 2
        act = torch.empty(states.shape[0], dtype=torch.int16, device=dev)
        for m in (0, 1):
            sel = mover == m
            if bool(sel.any()):
                counts, actions, _ = searchers[m].search(states[sel].contiguous())
                pol = pv_mcts.policy_from_counts(counts, temperature)
                pick = torch.multinomial(pol.float(), 1, generator=gen)
                act[sel] = torch.gather(actions, 1, pick).squeeze(1)
        if details:
            for g, a in zip(gid.tolist(), act.tolist()):
                log[g].append(int(a))
        states, term = gl.next_batch(states, act)
        ply += 1
        done = term != 0
        if bool(done.any()):
            lose = (term[done] & 1) != 0
            # ended state's player to move has lost; it is the first player iff ply is even
            fp = torch.where(lose, torch.full_like(lose, 0.0 if ply % 2 == 0 else 1.0, dtype=torch.float64),
                             torch.full_like(lose, 0.5, dtype=torch.float64))
            points[gid[done]] = fp
        states, gid = states[~done].contiguous(), gid[~done]
    g = torch.arange(num_games, device=dev)
    model0_points = torch.where(g % 2 == 0, points, 1.0 - points)
    total = float(model0_points.sum().item())
    return (total, points.tolist(), log) if details else total


def evaluate_network(model_dir=PV_NETWORK_PATH, num_games=EN_GAME_COUNT, sims=None):
    model0 = GNNNetwork()
    model0.prep_for_inference(os.path.join(model_dir, 'latest.pth'))
    model1 = GNNNetwork()
    model1.prep_for_inference(os.path.join(model_dir, 'best.pth'))
    total_point = play_matches(model0, model1, num_games, EN_TEMPERATURE, sims)
    average_point = total_point / num_games
    print('Average points of latest model against current best:', average_point)
    del model0, model1
    torch.cuda.empty_cache()
    if average_point > 0.5:
        update_best_player(model_dir)
        return True
    return False


if __name__ == '__main__':
    evaluate_network()
