"""Self-play -- drop-in for reference self_play.py, with all games of a run played in lock-step on
the GPU (game-level sharding across ranks needs no communication).

History format is the reference's (self_play.py:51-54,63-66): a list of
``[[player, enemy, walls], policy (209 floats), value]`` pickled to ``./data/<timestamp>.history``."""
import os
import pickle
from datetime import datetime

import numpy as np
import torch

from . import _lib
from . import game_logic as gl
from . import pv_mcts
from .constants import PV_NETWORK_PATH
from .pv_network_gnn import GNNNetwork, POLICY_OUTPUT_SIZE
from .positions import start_states

# Parameters
SP_GAME_COUNT = 50  # Number of games for self-play (self_play.py:19)
SP_TEMPERATURE = 1.0  # Temperature parameter for Boltzmann distribution (self_play.py:20)


def first_player_value(ended_state):
    """1: First player wins, -1: First player loses, 0: Draw (self_play.py:22-27)."""
    if ended_state.is_lose():
        return -1 if ended_state.is_first_player() else 1
    return 0


def write_data(history, directory='./data/'):
    """Save training data to a file (self_play.py:30-37)."""
    now = datetime.now()
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, '{:04}{:02}{:02}{:02}{:02}{:02}.history'.format(
        now.year, now.month, now.day, now.hour, now.minute, now.second))
    with open(path, mode='wb') as f:
        pickle.dump(history, f)
    return path


@torch.no_grad()
def play_batch_device(model, num_games, device=None, sims=None, temperature=SP_TEMPERATURE, seed=None, max_plies=None,
                      policy_dtype=torch.float64):
    """Execute `num_games` self-play games in lock-step (self_play.py:40-68 per game) and keep the record on the device:
    dict(states uint8[T,32] packed, policy [T,209] search policies over all actions (self_play.py:51-54), value f32[T]
    back-filled game results seen from the player to move (self_play.py:63-66), game int64[T], ply int64[T],
    flags uint8[G] (bit 0 = the player to move at the end has lost, bit 1 = draw), plies int64[G], sims int = simulations run).

    Per ply: one search (200 graph-replayed simulation steps in a handful of launches) and ONE library call,
    aq_selfplay_advance -- search policy, its dense 209-wide record, the sampled move, the next states and the compaction of the
    games still running -- then one synchronisation that reads the search's status flags and the number of survivors together.
    A game's moves are drawn from a hash of (seed, game, ply): they do not depend on the other games of the batch."""
    if policy_dtype not in (torch.float64, torch.float32):
        raise ValueError("policy_dtype must be torch.float64 or torch.float32")
    dev = gl._dev(device)
    L, P = _lib.load(), _lib.ptr
    seed = (int(seed) if seed is not None else int(torch.seed())) & (2 ** 64 - 1)
    sims = sims or pv_mcts.PV_EVALUATE_COUNT
    mcts = pv_mcts.searcher_for(model, sims, dev)
    states = start_states(num_games, dev)
    game_id = torch.arange(num_games, device=dev)
    rec_state, rec_policy, rec_game, sizes = [], [], [], []
    final_flags = torch.zeros(max(num_games, 1), dtype=torch.uint8, device=dev)
    final_plies = torch.zeros(max(num_games, 1), dtype=torch.int64, device=dev)
    ws = torch.empty((max(1, L.aq_selfplay_ws_bytes(num_games)),), dtype=torch.uint8, device=dev)
    alive = torch.zeros((1,), dtype=torch.int32, device=dev)
    ply, sims_run = 0, 0
    with torch.cuda.device(dev):
        while states.shape[0] > 0 and (max_plies is None or ply < max_plies):
            G = states.shape[0]
            counts, actions, n, status = mcts.search_async(states)
            sims_run += G * sims
            dense = torch.empty((G, POLICY_OUTPUT_SIZE), dtype=policy_dtype, device=dev)
            nxt = torch.empty_like(states)
            nxt_id = torch.empty_like(game_id)
            _lib.check(L.aq_selfplay_advance(P(states), P(counts), P(actions), P(n), P(game_id), G, float(temperature), seed, ply,
                                             P(dense), int(policy_dtype == torch.float64), None, P(nxt), P(nxt_id), P(final_flags),
                                             P(final_plies), P(alive), P(ws), _lib.stream_ptr(dev)), "aq_selfplay_advance")
            flags = torch.cat([status, alive]).tolist()      # the ply's only synchronisation
            mcts.check_status(flags)
            rec_state.append(states)
            rec_policy.append(dense)
            rec_game.append(game_id)
            sizes.append(G)
            ply += 1
            states, game_id = nxt[:flags[2]], nxt_id[:flags[2]]
    if not sizes:
        z = torch.zeros((0,), dtype=torch.int64, device=dev)
        return {"states": states, "policy": torch.zeros((0, POLICY_OUTPUT_SIZE), dtype=policy_dtype, device=dev),
                "value": torch.zeros((0,), dtype=torch.float32, device=dev), "game": z, "ply": z, "flags": final_flags[:num_games],
                "plies": final_plies[:num_games], "sims": 0}
    game = torch.cat(rec_game)
    plyv = torch.repeat_interleave(torch.arange(len(sizes), device=dev), torch.tensor(sizes, device=dev))
    final_flags, final_plies = final_flags[:num_games], final_plies[:num_games]
    # first_player_value of the ended state (self_play.py:22-27): the player to move there has lost; the value target of a
    # position alternates in sign with the ply (self_play.py:63-66); unfinished games (max_plies) count as draws
    lost = (final_flags[game] & 1) != 0
    fpv = torch.where(lost, torch.where(final_plies[game] % 2 == 0, -1.0, 1.0), 0.0)
    value = torch.where(plyv % 2 == 0, fpv, -fpv).to(torch.float32)
    return {"states": torch.cat(rec_state), "policy": torch.cat(rec_policy), "value": value, "game": game, "ply": plyv,
            "flags": final_flags, "plies": final_plies, "sims": sims_run}


def play_batch(model, num_games, device=None, sims=None, temperature=SP_TEMPERATURE, seed=None, max_plies=None):
    """play_batch_device + conversion to the reference's history: a list of [[player, enemy, walls], policy (209 floats), value]
    (self_play.py:51-54, 63-66), games in order, plies in order within a game.  Returns (history, dict with per-game results)."""
    rec = play_batch_device(model, num_games, device, sims, temperature, seed, max_plies)
    rows, _ = gl.unpack_rows(rec["states"])
    rows = rows.cpu().numpy()
    pols = rec["policy"].cpu().numpy()
    vals = rec["value"].cpu().numpy()
    gids = rec["game"].cpu().numpy()
    order = np.argsort(gids, kind='stable')                              # plies stay in order within a game
    history = []
    for i in order:
        r = rows[i]
        history.append([[r[0:2].tolist(), r[2:4].tolist(), r[4:].tolist()], pols[i].tolist(), int(vals[i])])
    return history, {"flags": rec["flags"].cpu().numpy(), "plies": rec["plies"].cpu().numpy()}


def play(model, device=None):
    """Execute one self-play game (self_play.py:40-68)."""
    history, _ = play_batch(model, 1, device)
    return history


def self_play(game_count=None, model_path=None, data_dir='./data/', rank=0, world_size=1, sims=None, seed=None):
    """Perform self-play games and save the training data (self_play.py:71-98).  With
    world_size > 1 each rank plays its share of the games on its own GPU and rank 0 writes the file."""
    game_count = SP_GAME_COUNT if game_count is None else game_count
    model = GNNNetwork()
    model.prep_for_inference(model_path=model_path or (PV_NETWORK_PATH + 'best.pth'))
    mine = game_count // world_size + (1 if rank < game_count % world_size else 0)
    history, _ = play_batch(model, mine, sims=sims, seed=None if seed is None else seed + rank) if mine else ([], None)
    if world_size > 1:
        import torch.distributed as dist
        gathered = [None] * world_size if rank == 0 else None
        dist.gather_object(history, gathered, dst=0)
        if rank == 0:
            history = [h for part in gathered for h in part]
    path = None
    if rank == 0:
        path = write_data(history, data_dir)
        print(f'Self-play: {game_count} games, {len(history)} positions -> {path}')
    del model
    torch.cuda.empty_cache()
    return path


if __name__ == '__main__':
    self_play()
