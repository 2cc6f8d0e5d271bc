"""Synthetic legal positions generated on the GPU with the product's own kernels (K1 legal mask +
K9 state_next): G games are played in lock-step from the start position with the move mix of
SURVEY.md section 8d (probability 0.5 a uniformly random legal wall action if any, else a uniformly
random legal pawn action) and every visited non-terminal position is harvested.  Used by bench.py
and the size-independent GPU tests; the oracle is NOT involved."""
import torch

from . import game_logic as gl


def start_states(G, device=None):
    rows = torch.zeros((G, 68), dtype=torch.uint8)
    rows[:, 0] = 76
    rows[:, 1] = 10
    rows[:, 2] = 76
    rows[:, 3] = 10
    plies = torch.zeros((G,), dtype=torch.int16)
    return gl.pack_rows(rows, plies, device)


@torch.no_grad()
def random_positions(n, seed=1, games=8192, device=None, max_plies=116):
    """-> packed uint8[n,32] CUDA tensor of reachable, non-terminal positions (trajectory harvest)."""
    dev = gl._dev(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    out, total = [], 0
    while total < n:
        packed = start_states(games, dev)
        for _ in range(max_plies):
            if packed.shape[0] == 0:
                break
            mask, _ = gl.legal_mask_batch(packed)
            dense = gl.mask_to_dense(mask)
            has_any = dense.any(dim=1)
            cur, dense = packed[has_any], dense[has_any]
            if cur.shape[0] == 0:
                break
            out.append(cur)
            total += cur.shape[0]
            if total >= n:
                break
            walls = dense.clone()
            walls[:, : gl.NUM_SQUARES] = False
            pawns = dense.clone()
            pawns[:, gl.NUM_SQUARES:] = False
            coin = torch.rand((cur.shape[0],), device=dev, generator=gen) < 0.5
            use_wall = walls.any(dim=1) & (coin | ~pawns.any(dim=1))
            pick = torch.where(use_wall.unsqueeze(1), walls, pawns).float()
            action = torch.multinomial(pick, 1, generator=gen).squeeze(1).to(torch.int16)
            nxt, term = gl.next_batch(cur, action)
            packed = nxt[term == 0].contiguous()
    return torch.cat(out, 0)[:n].contiguous()


def mixed_batches(nb, B, seed=1, device=None):
    """nb batches of B positions with the same mix of game phases (plies 0..~31: full wall racks down to endgames):
    one trajectory harvest of nb*B positions from nb*B/32 lock-step games, shuffled with a fixed permutation.
    The workload of bench.py and of the committed ncu captures."""
    dev = gl._dev(device)
    allpos = random_positions(nb * B, seed=seed, games=max(64, nb * B // 32), device=dev)
    g = torch.Generator(device="cpu")
    g.manual_seed(1234 + seed)
    allpos = allpos[torch.randperm(nb * B, generator=g).to(dev)].contiguous()
    return allpos, [allpos[i * B:(i + 1) * B].contiguous() for i in range(nb)]
