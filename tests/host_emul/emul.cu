// Host-side emulation of the legal-mask kernel's per-state algorithm, compiled from the SAME
// device header (aq_common.cuh is __host__ __device__).  Test-only: lets the CPU suite check the
// bitboard restatement (flood fill with pawn rules, bit-parallel gate) against the golden
// fixtures before any GPU time is spent.  The warp-level plumbing of the real kernel
// (compaction, OR-reduction) is covered by the -m gpu tests.
#include "../../alphaquoridorgnn_b200/csrc/aq_common.cuh"
using namespace aq;

extern "C" void emul_pack(const uint8_t *rows, const int16_t *plies, long long B, AqState *out) {
    for (long long b = 0; b < B; ++b) {
        const uint8_t *r = rows + 68 * b;
        AqState s{};
        for (int i = 0; i < 64; ++i) {
            if (r[4 + i] == 1) s.hwalls |= 1ull << i;
            if (r[4 + i] == 2) s.vwalls |= 1ull << i;
        }
        s.ppos = r[0]; s.pwalls = r[1]; s.epos = r[2]; s.ewalls = r[3];
        s.plies = plies ? (uint16_t)plies[b] : 0;
        out[b] = s;
    }
}

extern "C" void emul_legal_mask(const AqState *states, long long B, uint32_t *mask, uint8_t *pawn) {
    for (long long b = 0; b < B; ++b) {
        const AqState s = states[b];
        const Open base = open_from_walls(s.hwalls, s.vwalls);
        const int me = s.ppos, en = 80 - (int)s.epos;
        u64 legalH = 0, legalV = 0;
        if (s.pwalls > 0) {
            const WallSets ws = wall_sets(s.hwalls, s.vwalls);
            legalH = ws.freeH;
            legalV = ws.freeV;
            // same decision structure as legal_mask_kernel: witness paths first, a real search only for the
            // player whose witness the candidate cuts
            const PathCuts pm_ = find_path_cuts(base, me, en, goal_row0());
            const PathCuts pe_ = find_path_cuts(base, en, me, goal_row8());
            for (int slot = 0; slot < 64; ++slot)
                for (int orient = 1; orient <= 2; ++orient) {
                    const u64 need = orient == 1 ? ws.needH : ws.needV;
                    if (!((need >> slot) & 1)) continue;
                    const u64 cm = orient == 1 ? pm_.cutH : pm_.cutV, ce = orient == 1 ? pe_.cutH : pe_.cutV;
                    Open o = base;
                    add_wall(o, orient, slot);
                    bool ok = true;
                    if (!pm_.exists || ((cm >> slot) & 1)) ok = ok && reaches(o, me, en, goal_row0());
                    if (!pe_.exists || ((ce >> slot) & 1)) ok = ok && reaches(o, en, me, goal_row8());
                    if (ok) {
                        if (orient == 1) legalH |= 1ull << slot; else legalV |= 1ull << slot;
                    }
                }
        }
        uint8_t pm[8] = {0, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0, 0};
        const int n = pawn_moves(base, me, en, pm + 1);
        pm[0] = (uint8_t)n;
        u128 lo = 0;
        for (int k = 0; k < n; ++k) lo |= (u128)1 << pm[1 + k];
        lo |= (u128)legalH << 81;
        const u128 hi = (u128)(legalH >> 47) | ((u128)legalV << 17);
        uint32_t *m = mask + 8 * b;
        for (int k = 0; k < 4; ++k) { m[k] = (uint32_t)(lo >> (32 * k)); m[4 + k] = (uint32_t)(hi >> (32 * k)); }
        for (int k = 0; k < 8; ++k) pawn[8 * b + k] = pm[k];
    }
}

extern "C" void emul_open_mask(const AqState *states, long long B, uint8_t *open_mask) {
    for (long long b = 0; b < B; ++b) {
        const Open o = open_from_walls(states[b].hwalls, states[b].vwalls);
        for (int v = 0; v < 81; ++v)
            open_mask[b * 81 + v] = (uint8_t)((int)has(o.up, v) | ((int)has(o.down, v) << 1) | ((int)has(o.left, v) << 2) | ((int)has(o.right, v) << 3));
    }
}

// shortest_paths_kernel's per-state computation (agents.heuristic_eval, agents.py:22-54)
extern "C" void emul_shortest_paths(const AqState *states, long long B, int16_t *dist) {
    for (long long b = 0; b < B; ++b) {
        const AqState s = states[b];
        const Open o = open_from_walls(s.hwalls, s.vwalls);
        const int me = s.ppos, en = 80 - (int)s.epos;
        dist[2 * b] = (int16_t)path_length(o, me, en, goal_row0());
        dist[2 * b + 1] = (int16_t)path_length(o, en, me, goal_row8());
    }
}
